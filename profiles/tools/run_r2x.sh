set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300  2>&1 | tail -15 > gpurun_out/r2x_pytest_quick.log
cat gpurun_out/r2x_pytest_quick.log
SPLISER_TIMING=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-bam --no-variants --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
grep compact gpurun_out/r2x_bench.err | tail -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","launches_per_step","parity_checked"): print(k, d.get(k))
print(d["roofline_path"]["kernel_ms"]); print(json.dumps(d["e2e"])[:400])
PY
