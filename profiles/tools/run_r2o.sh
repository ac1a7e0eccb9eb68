set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -8 > gpurun_out/r2o_pytest.log
cat gpurun_out/r2o_pytest.log
timeout 600 python bench.py --workload c4 --reads 200000 --gpus 1 --steps 2 > gpurun_out/r2o_c4.json 2> gpurun_out/r2o_c4.err
tail -3 gpurun_out/r2o_c4.err; cat gpurun_out/r2o_c4.json | head -c 2500
timeout 600 python profiles/tools/quick_time.py c2 40000000 fused > gpurun_out/r2o_quick.json 2> gpurun_out/r2o_quick.err
cat gpurun_out/r2o_quick.json
