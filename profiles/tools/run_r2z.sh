set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "appendix or golden_process_fuzz or live_fuzz or synthetic_workload or stabbing" 2>&1 | tail -8 > gpurun_out/r2z_pytest_quick.log
cat gpurun_out/r2z_pytest_quick.log
timeout 300 python profiles/tools/quick_time.py c2 40000000 fused > gpurun_out/r2z_quick.json 2> gpurun_out/r2z_quick.err
cat gpurun_out/r2z_quick.json; tail -3 gpurun_out/r2z_quick.err
