set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2k_pytest.log
cat gpurun_out/r2k_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","launches_per_step","parity_checked"): print(k, d.get(k))
print(d["roofline_path"]["kernel_ms"]); print(json.dumps(d["e2e"])[:900]); print(d.get("bam_e2e",{}).get("ms_per_step"), d.get("cli_e2e",{}).get("ms_per_step"))
PY
