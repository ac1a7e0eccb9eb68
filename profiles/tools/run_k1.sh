# K1 development loop: parity slice, then resident timing on configs[1] with the phase timeline of the cooperative kernel
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "appendix or golden_process_fuzz or synthetic_workload or graph_builders or graph_builds or tiles or dirty or full_size" 2>&1 | tail -5
SPLISER_K1_STAMPS=1 timeout 300 python profiles/tools/quick_time.py c2 40000000 fused 2>gpurun_out/k1_coop.err | tee gpurun_out/k1_coop.json
grep stamps gpurun_out/k1_coop.err
