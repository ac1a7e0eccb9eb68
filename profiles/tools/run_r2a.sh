set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_pytest.log
timeout 600 python profiles/tools/quick_time.py c2 40000000 > gpurun_out/r2a_quick.json 2> gpurun_out/r2a_quick.err
cat gpurun_out/r2a_quick.json; tail -5 gpurun_out/r2a_quick.err
