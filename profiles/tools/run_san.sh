set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bash profiles/tools/sanitize.sh > gpurun_out/san_final.log 2>&1
grep "===\|ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/san_final.log gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_synccheck.log gpurun_out/sanitize_initcheck.log | tail -20
grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitize_racecheck.log
grep "Race reported" gpurun_out/sanitize_racecheck.log | grep -v inflate_member | head -5
