set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "not full_size and not at_scale" 2>&1 | tail -4
SPLISER_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-bam --no-variants --no-cpu-baseline --e2e-steps 5 > gpurun_out/e2e_bench.json 2> gpurun_out/e2e_bench.err
grep compact gpurun_out/e2e_bench.err | tail -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/e2e_bench.json').read().strip().splitlines()[-1])
e=d["e2e"]; print(d["value"], e["ms_per_step"], e["breakdown_ms"], e["packed_view"]["ms_per_step"], e["plain_view"]["ms_per_step"])
PY
