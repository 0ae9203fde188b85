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
    # training: ONE all-reduce over the flat gradient buffer every parameter's .grad is a view of (section 8e)
    from rsis_b200.autograd import GradBucket
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)),
              torch.nn.Parameter(torch.randn(2, 2, 3), requires_grad=False)]
    bucket = GradBucket(params)
    assert bucket.flat.numel() == 22 and params[2].grad is None
    bucket.zero()
    loss = sum(((rank + 1.0) * p * p).sum() for p in params[:2])   # rank-dependent gradients
    loss.backward()                                                # autograd accumulates into the views in place
    assert params[0].grad.data_ptr() == bucket.flat.data_ptr()
    bucket.all_reduce()                                            # sum over ranks, / world
    want = torch.cat([(2 * 1.5 * p.detach()).reshape(-1) for p in params[:2]])   # mean of 2*(rank+1)*p over ranks
    assert torch.allclose(bucket.flat, want, atol=1e-6), (bucket.flat, want)
    for p_ in params[:2]:
        p_.grad = None                                            # an optimiser's zero_grad(set_to_none=True)
    bucket.zero()                                                  # ... re-attaches the views
    assert params[1].grad is not None and float(bucket.flat.abs().sum()) == 0.0
    # masked class / stop losses (SURVEY.md section 8e): the reference's DataParallel criteria return the un-reduced
    # selected vectors, averaged GLOBALLY (train.py:161,168).  With gradients averaged over ranks each rank contributes
    # local_sum * world / n_valid_global -- rsis_b200.objectives.masked_mean.
    from rsis_b200.objectives import masked_mean
    gen = torch.Generator().manual_seed(3)
    costs_all = torch.rand(10, generator=gen)
    sel_all = torch.tensor([1, 0, 1, 1, 0, 0, 1, 0, 0, 1], dtype=torch.bool)   # 3 valid rows on rank 0, 2 on rank 1
    lo, hi = rd.shard_range(10, rank, world)
    c = torch.where(sel_all[lo:hi], costs_all[lo:hi], torch.zeros(hi - lo))
    sc = torch.stack([c.sum(), sel_all[lo:hi].float().sum()])
    mine = masked_mean(c, sc)
    both = mine.clone()
    dist.all_reduce(both)
    want = costs_all[sel_all].mean()
    assert abs(float(both / world) - float(want)) < 1e-6, (both, want)       # DDP-style average of per-rank losses
    assert abs(float(mine) - float(c.sum() * 2 / 5)) < 1e-6
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
