set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2f_pytest.log
cat gpurun_out/r2f_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -c 3000 gpurun_out/r2f_bench.err | tail -20
cat gpurun_out/r2f_bench.json | head -c 6000
timeout 1500 bash profiles/tools/sanitize.sh > gpurun_out/r2f_sanitize.log 2>&1
grep -E "===|ERROR SUMMARY|passed|failed" gpurun_out/r2f_sanitize.log
