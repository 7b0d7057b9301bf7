"""CPU test (gloo, world_size 2) of the N>1 host logic of bench.py: rank bootstrap from the torchrun environment,
barrier + max-over-ranks timing, whole-job aggregation, and the reference arm's "rank 0 only" rule."""
import json
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, {root!r})
    import bench
    rank, world, local, dist = bench.dist_setup(2, backend="gloo")
    assert world == 2 and dist is not None
    mine = 10.0 if rank == 0 else 25.0                       # pretend decode-segment milliseconds
    worst = bench.barrier_max(dist, local, mine)
    val = bench.aggregate_tokens_per_s(world, steps=2, tokens_per_step=496, ms_per_rank_max=worst)
    dist.barrier()
    print(json.dumps({{"rank": rank, "worst": worst, "value": val}}))
    dist.destroy_process_group()
""")


def test_gloo_world2_max_over_ranks_and_aggregation(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    res = [json.loads(o.strip().splitlines()[-1]) for o, _ in outs]
    assert all(r["worst"] == 25.0 for r in res)                          # max over ranks, identical everywhere
    assert all(abs(r["value"] - 2 * 2 * 496 / 0.025) < 1e-6 for r in res)  # all ranks' tokens over the slowest rank's time


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference under torchrun: rank 0 prints the line, other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29632")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "toy-qwen3-q80", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env["RANK"] = "0"; env["LOCAL_RANK"] = "0"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "toy-qwen3-q80", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0
