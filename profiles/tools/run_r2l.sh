set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -15 > gpurun_out/r2l_pytest.log
cat gpurun_out/r2l_pytest.log
