set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 2>&1 | tail -8 > gpurun_out/r3d_pytest.log
cat gpurun_out/r3d_pytest.log
timeout 300 python profiles/tools/quick_time.py c2 40000000 fused > gpurun_out/r3d_quick.json 2> gpurun_out/r3d_quick.err
cat gpurun_out/r3d_quick.json
timeout 300 python profiles/tools/quick_time.py c3 50000000 fused > gpurun_out/r3d_quick_c3.json 2> gpurun_out/r3d_quick_c3.err
cat gpurun_out/r3d_quick_c3.json
