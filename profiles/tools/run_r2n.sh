set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -5 > gpurun_out/r2n_pytest.log
cat gpurun_out/r2n_pytest.log
timeout 600 python bench.py --profile --steps 3 --warmup 3 > /dev/null 2>&1   # warm the cache
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_launches_c2.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/r2n_ncu_c2.log 2>&1
timeout 600 python bench.py --workload c3 --reads 100000000 --profile --steps 3 --warmup 3 > gpurun_out/r2n_c3_profile.json 2> gpurun_out/r2n_c3.err
cat gpurun_out/r2n_c3_profile.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_launches_c3.csv python bench.py --workload c3 --reads 100000000 --profile --steps 2 --warmup 3 > gpurun_out/r2n_ncu_c3.log 2>&1
ls -la gpurun_out | tail -8
