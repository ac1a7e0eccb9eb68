set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
free -g | head -2; nproc
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --reads 40000000 > gpurun_out/r2g_bench2.json 2> gpurun_out/r2g_bench2.err
tail -5 gpurun_out/r2g_bench2.err
cat gpurun_out/r2g_bench2.json | head -c 5000
timeout 600 python bench.py --impl reference --gpus 2 --steps 3 --warmup 1 --reads 40000000 > gpurun_out/r2g_ref2.json 2> gpurun_out/r2g_ref2.err
cat gpurun_out/r2g_ref2.json | head -c 1500
