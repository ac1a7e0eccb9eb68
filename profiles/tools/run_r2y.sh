set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
tail -3 gpurun_out/r2y_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2y_launches_c2.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/r2y_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_count_fused|k_hot_items" -s 4 -c 2 -f -o gpurun_out/r2y_fused python profiles/tools/quick_time.py c2 40000000 fused > gpurun_out/r2y_ncu.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2y_reference.json 2> gpurun_out/r2y_reference.err
cat gpurun_out/r2y_reference.json | cut -c1-400
ls -la gpurun_out | tail -8
