#!/bin/bash
# compute-sanitizer passes over a small slice of the GPU parity tests (one B200, under gpurun):
#   gpurun --timeout 1500 -- 'bash profiles/tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# memcheck (out-of-bounds / misaligned accesses, leaks of device memory), racecheck (shared-memory hazards in the
# warp-specialised kernels: the TMA rings of k_count_fused and K3, the per-warp hot lists, the junction kernels' work lists, the inflate tables), synccheck (barrier misuse),
# initcheck (reads of uninitialised device memory).  The selected tests cover every kernel: golden fuzz (all modes, dirty
# regime), a synthetic sample through the records path and through the device BAM ingest, the re-count entry point.
set -u
SEL='test_appendix_a_known_answers or test_empty_and_ragged_inputs or test_bam_ingest_on_the_device or test_golden_combine_recount or test_stabbing_variant_equals_difference_array_variant or test_compact_view_equals_plain_view or test_junction_extraction'
export SPLISER_SANITIZE_SMALL=1
for tool in memcheck racecheck synccheck initcheck; do
    echo "=== compute-sanitizer --tool $tool"
    timeout 1200 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
        python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/sanitize_$tool.log 2>&1
    echo "=== $tool exit code: $?"
    tail -n 25 gpurun_out/sanitize_$tool.log
done
