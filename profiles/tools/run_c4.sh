set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --workload c4 --gpus 1 --steps 2 --warmup 1 > gpurun_out/c4_1gpu.json 2> gpurun_out/c4_1gpu.err
tail -2 gpurun_out/c4_1gpu.err
timeout 900 python bench.py --workload c4 --gpus 2 --steps 2 --warmup 1 > gpurun_out/c4_2gpu.json 2> gpurun_out/c4_2gpu.err
tail -2 gpurun_out/c4_2gpu.err
python - <<'PY'
import json
for f in ("c4_1gpu","c4_2gpu"):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["n_gpus"], d["process_per_sample_ms"], d["config"]["combined_rows"])
PY
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/scale_2.json 2> gpurun_out/scale_2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/scale_2.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","parity_checked"): print(k, d.get(k))
print(d["roofline_path"]["kernel_ms"]); print(d["e2e"]["value"], d["e2e"]["ms_per_step"]); print(json.dumps(d.get("strong_scaling"))[:300])
PY
