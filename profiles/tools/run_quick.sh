# development loop on one B200 under gpurun: a slice of the parity tests, then the resident timing of the fused path
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "appendix or golden_process_fuzz or synthetic_workload" 2>&1 | tail -3
timeout 300 python profiles/tools/quick_time.py c2 40000000 fused 2>/dev/null
