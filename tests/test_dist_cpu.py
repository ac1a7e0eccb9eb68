"""N > 1 host logic on the CPU: two gloo ranks exercise the rank helpers bench.py uses (barrier, max / sum
over ranks, per-rank workloads, tile and sample sharding)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gloo_ranks(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent('''
        import sys
        sys.path.insert(0, %r)
        import numpy as np
        from spliser_b200 import synth
        from spliser_b200.dist import Ranks, samples_of, tile_of
        import bench
        r = Ranks("gloo")
        assert r.world == 2
        cfg, _ = bench.workload_config("small", 4000, r.rank)
        w = synth.generate(cfg)
        r.barrier()
        total = r.sum(len(w.records))
        slow = r.max(10.0 + r.rank)
        chk = r.sum(float(w.records.pos[:100].sum()))
        mine = float(w.records.pos[:100].sum())
        assert total == 8000 and slow == 11.0
        assert chk != 2 * mine            # the two ranks count different samples (per-rank seed)
        lo, hi = tile_of(r.rank, r.world, 101)
        assert r.sum(hi - lo) == 101
        assert samples_of(r.rank, r.world, 5) == ([0, 2, 4] if r.rank == 0 else [1, 3])
        r.close()
        print("rank", r.rank, "ok")
    ''' % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


def test_bench_reference_arm_prints_the_contract_line(tmp_path):
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a small workload: ONE JSON line with the
    keys of the bench contract; under torchrun (N > 1) rank 0 alone prints it and the other rank exits 0 without work."""
    import json
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", SPLISER_BENCH_CACHE=str(tmp_path))
    tail = ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--reads", "120000", "--cpu-sample", "40000"]
    one = subprocess.run([sys.executable] + tail, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert one.returncode == 0, one.stderr[-2000:]
    lines = [ln for ln in one.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "reads/s" and d["value"] > 0 and d["warmup"] >= 3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29519"] + tail[:1] + ["--gpus", "2"] + tail[1:], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert two.returncode == 0, two.stderr[-2000:]
    lines = [ln for ln in two.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2
