# the whole GPU suite + smoke on the final .so
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 160 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -4 > gpurun_out/f3_pytest.log
cat gpurun_out/f3_pytest.log
