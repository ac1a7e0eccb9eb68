set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2j_pytest.log
cat gpurun_out/r2j_pytest.log
timeout 600 python profiles/tools/quick_time.py c2 40000000 fused > gpurun_out/r2j_quick.json 2> gpurun_out/r2j_quick.err
cat gpurun_out/r2j_quick.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_count_fused -s 2 -c 1 -f -o gpurun_out/r2j_fused python profiles/tools/quick_time.py c2 40000000 fused > gpurun_out/r2j_ncu.log 2>&1
