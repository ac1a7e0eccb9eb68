set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "compact or packed or empty_and_ragged" 2>&1 | tail -12 > gpurun_out/r3e_pytest.log
cat gpurun_out/r3e_pytest.log
