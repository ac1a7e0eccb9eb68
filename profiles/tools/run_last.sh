set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SPLISER_K1_STAMPS=1 timeout 25 python profiles/tools/k1_grch38_stamps.py 8000000 > gpurun_out/k1_grch38_hint.json 2> gpurun_out/k1_grch38_hint.err
cat gpurun_out/k1_grch38_hint.json; grep stamps gpurun_out/k1_grch38_hint.err
timeout 12 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 10 -k "appendix or golden_process_fuzz or graph_builders or dirty" 2>&1 | tail -2
