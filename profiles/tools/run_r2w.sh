set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python profiles/tools/e2e_profile.py 40000000 2 > gpurun_out/r2w_e2e.json 2> gpurun_out/r2w_e2e.err
cat gpurun_out/r2w_e2e.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2w_launches_e2e.csv python profiles/tools/e2e_profile.py 40000000 1 > gpurun_out/r2w_ncu.log 2>&1
tail -2 gpurun_out/r2w_ncu.log
