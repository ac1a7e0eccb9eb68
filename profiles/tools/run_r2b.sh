set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# warm the workload cache (no profiler)
timeout 600 python profiles/tools/quick_time.py c2 40000000 fused > gpurun_out/r2b_quick.json 2> gpurun_out/r2b_quick.err
cat gpurun_out/r2b_quick.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_count_fused -s 2 -c 1 -f -o gpurun_out/r2b_fused python profiles/tools/quick_time.py c2 40000000 fused > gpurun_out/r2b_ncu.log 2>&1
tail -5 gpurun_out/r2b_ncu.log
ls -la gpurun_out/
