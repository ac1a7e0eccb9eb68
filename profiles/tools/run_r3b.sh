set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 400 -k "full_size_configs1 or configs2_tile_at_scale or compact" 2>&1 | tail -8 > gpurun_out/r3b_pytest.log
cat gpurun_out/r3b_pytest.log
bash profiles/tools/sanitize.sh > gpurun_out/r3b_sanitize.log 2>&1
grep "===\|ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/r3b_sanitize.log gpurun_out/sanitize_*.log | tail -30
