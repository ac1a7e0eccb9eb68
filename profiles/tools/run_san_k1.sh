# compute-sanitizer over the cooperative K1 (k_gb_coop: grid-wide barriers, per-warp shared-memory counters and lists): the appendix
# known-answer cases + the junction extraction (which shares the sort / scan kernels), all four tools
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SEL='test_appendix_a_known_answers or test_junction_extraction'
export SPLISER_SANITIZE_SMALL=1
for tool in memcheck synccheck racecheck initcheck; do
    echo "=== compute-sanitizer --tool $tool"
    timeout 75 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
        python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/sanitize_k1_$tool.log 2>&1
    echo "=== $tool exit code: $?"
    grep "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed\|k_gb_coop" gpurun_out/sanitize_k1_$tool.log | tail -6
done
