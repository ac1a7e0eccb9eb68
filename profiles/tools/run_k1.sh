# K1 as one cooperative kernel against the launch-per-phase build: parity slice, then resident timing of both on configs[1]
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "appendix or golden_process_fuzz or synthetic_workload or graph_builders or tiles or dirty" 2>&1 | tail -5
SPLISER_K1_STAMPS=1 timeout 300 python profiles/tools/quick_time.py c2 40000000 fused 2>gpurun_out/k1_coop2.err | tee gpurun_out/k1_coop2.json
grep stamps gpurun_out/k1_coop2.err
SPLISER_K1_STAMPS=1 SPLISER_K1_CTAS_PER_SM=1 timeout 300 python profiles/tools/quick_time.py c2 40000000 fused 2>gpurun_out/k1_coop1.err | tee gpurun_out/k1_coop1.json
grep stamps gpurun_out/k1_coop1.err
