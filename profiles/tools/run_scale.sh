set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4}
date
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${N}.json 2> gpurun_out/scale_${N}.err
date
tail -3 gpurun_out/scale_${N}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_${N}.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","launches_per_step","parity_checked","scaling"): print(k, d.get(k))
print(d["roofline_path"]["kernel_ms"]); print(d["e2e"]["value"], d["e2e"]["ms_per_step"]); print(json.dumps(d.get("strong_scaling"))[:400]); print(d["parity"]["tiles_digest"], d["parity"]["single_gpu_digest"], d["parity"]["oracle_subset"]["identical"])
PY
