set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
date
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/r3c_bench8.json 2> gpurun_out/r3c_bench8.err
date
tail -5 gpurun_out/r3c_bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3c_bench8.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","launches_per_step","parity_checked","scaling"): print(k, d.get(k))
print(d["roofline_path"]["kernel_ms"]); print(json.dumps(d["e2e"])[:500]); print(json.dumps(d.get("strong_scaling"))); print(json.dumps(d.get("parity"))[:900]); print(d["config"]["tile_graph"])
PY
