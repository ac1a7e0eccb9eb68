set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 200 -k "own_graph or tile" 2>&1 | tail -8 > gpurun_out/r2q_pytest.log
cat gpurun_out/r2q_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --reads 50000000 --steps 10 --warmup 3 > gpurun_out/r2q_bench2.json 2> gpurun_out/r2q_bench2.err
tail -5 gpurun_out/r2q_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench2.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","launches_per_step","parity_checked","scaling"): print(k, d.get(k))
print(d["roofline_path"]["kernel_ms"]); print(json.dumps(d["e2e"])[:300]); print(json.dumps(d.get("strong_scaling"))); print(json.dumps(d.get("parity"))[:900]); print(d["config"]["tile_graph"])
PY
