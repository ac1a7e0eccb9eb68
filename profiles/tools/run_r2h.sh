set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2h_pytest.log
cat gpurun_out/r2h_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 --no-bam > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","launches_per_step","parity_checked"): print(k, d.get(k))
print(d["roofline_path"]["kernel_ms"]); print(d["e2e"]["ms_per_step"], d["e2e"]["breakdown_ms"])
PY
timeout 1500 bash profiles/tools/sanitize.sh > gpurun_out/r2h_sanitize.log 2>&1
grep -E "===|ERROR SUMMARY|passed|failed|RACECHECK" gpurun_out/r2h_sanitize.log
