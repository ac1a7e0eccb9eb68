set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "bam or configs0" 2>&1 | tail -12 > gpurun_out/r3g_pytest.log
cat gpurun_out/r3g_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 --no-variants > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err
tail -3 gpurun_out/r3g_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3g_bench.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","parity_checked"): print(k, d.get(k))
print(d["roofline_path"]["kernel_ms"]); print(d["e2e"]["ms_per_step"]); print(d["bam_e2e"]); print(d.get("bam_e2e_seq")); print(d["cli_e2e"]["ms_per_step"])
PY
