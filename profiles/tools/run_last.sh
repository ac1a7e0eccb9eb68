# proof that the sanitizer runs go through k_gb_coop (the phase timeline is only printed by the cooperative path)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SPLISER_K1_STAMPS=1 SPLISER_SANITIZE_SMALL=1 timeout 75 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -s -m gpu -k "test_appendix_a_known_answers" -p no:cacheprovider > gpurun_out/san_k1_stamps.log 2>&1
grep -c "k1 stamps" gpurun_out/san_k1_stamps.log; grep "k1 stamps" gpurun_out/san_k1_stamps.log | head -2; tail -3 gpurun_out/san_k1_stamps.log
