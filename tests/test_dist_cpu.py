"""Host-side data-parallel plumbing on CPU: two `gloo` ranks (SURVEY.md section 8e -- the path shards per image, the
only cross-rank traffic is the rendezvous, barriers and the max-over-ranks timing reduction bench.py uses)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, os.environ["RSIS_ROOT"])
    import torch, torch.distributed as dist
    from rsis_b200 import dist as rd
    rank, local_rank, world = rd.init_from_env(backend="gloo")
    assert world == 2 and rank == local_rank and dist.get_backend() == "gloo"
    # every image of a global batch is owned by exactly one rank
    n = 13
    b, e = rd.shard_range(n, rank, world)
    owned = torch.zeros(n, dtype=torch.int64)
    owned[b:e] = 1
    dist.all_reduce(owned)
    assert bool((owned == 1).all()), owned
    # timing reductions: max over ranks (the step time bench.py reports), sum (units processed)
    rd.barrier()
    assert rd.max_over_ranks(1.0 + rank, device="cpu") == 2.0
    assert rd.sum_over_ranks(float(e - b), device="cpu") == float(n)
    rd.barrier()
    print(f"rank {rank} ok [{b},{e})")
""")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_gloo_ranks_shard_and_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), RSIS_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert f"rank {rank} ok" in out
