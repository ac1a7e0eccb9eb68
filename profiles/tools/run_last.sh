# K1 on a GRCh38-shaped table (few records: the table is what K1 sees) with the phase timeline of k_gb_coop, then a short parity slice.
# Run twice at the end of round 2: with the shipped bin-index fill (K1 0.243 ms, k_gb_coop 0.111 of it) and with the search-hint
# variant of k_gb_sb_fill (0.261 ms: slower, reverted) -- DESIGN section 9.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SPLISER_K1_STAMPS=1 timeout 25 python profiles/tools/k1_grch38_stamps.py 8000000 > gpurun_out/k1_grch38.json 2> gpurun_out/k1_grch38.err
cat gpurun_out/k1_grch38.json; grep stamps gpurun_out/k1_grch38.err
timeout 12 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 10 -k "appendix or golden_process_fuzz or graph_builders or dirty" 2>&1 | tail -2
