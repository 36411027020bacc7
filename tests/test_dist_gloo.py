"""N>1 host logic on CPU: world_size-2 gloo process group (127.0.0.1 rendezvous)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_views, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    from mvs_b200 import dist as D
    r, l, w = D.init("gloo")
    mine = D.shard_ref_views(n_views, r, w)
    D.barrier()
    slowest = D.max_over_ranks(10.0 + 5.0 * r)           # rank 1 is slower: 15.0
    counts = D.gather_counts(len(mine))
    red = D.reduce_scalars_to_rank0({"abs_depth_error": 1.0 + r, "thres2mm_error": 0.5 * (r + 1)})
    q.put((r, list(mine), slowest, counts, red))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [7, 8, 1])
def test_shard_and_timing_world2(n_views):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, q)) for r in range(world)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs: p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    all_views = res[0][1] + res[1][1]
    assert all_views == list(range(n_views))                          # disjoint, complete, ordered
    assert abs(len(res[0][1]) - len(res[1][1])) <= 1                  # balanced
    assert res[0][2] == res[1][2] == 15.0                              # max over ranks
    assert res[0][3] == res[1][3] == [len(res[0][1]), len(res[1][1])]
    assert res[0][4] == {"abs_depth_error": 1.5, "thres2mm_error": 0.75}   # mean on rank 0


def test_shard_single_process_and_errors():
    from mvs_b200 import dist as D
    assert list(D.shard_ref_views(5, 0, 1)) == [0, 1, 2, 3, 4]
    assert [len(D.shard_ref_views(10, r, 4)) for r in range(4)] == [3, 3, 2, 2]
    assert list(D.shard_ref_views(0, 0, 2)) == []
    with pytest.raises(ValueError):
        D.shard_ref_views(4, 2, 2)
    assert D.max_over_ranks(3.5) == 3.5 and D.gather_counts(4) == [4]


def _bucket_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mvs_b200.train import GradBucket
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    list(net.parameters())[1].grad = None                      # a parameter without gradient contributes zeros
    b = GradBucket(net.parameters())
    b.reduce(); b.wait()
    q.put((rank, [float(p.grad.flatten()[0]) for p in net.parameters()], b.numel))
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_world2():
    """GradBucket (the DDP role of CasMVSNet/train.py:367-372): one flat bucket, SUM all-reduce, mean over ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 431
    ps = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
    for rank, grads, numel in res:
        # mean over ranks of (rank+1)*(i+1) = 1.5*(i+1); the grad-less parameter averages (0 + 0) / 2 = 0
        assert grads == [1.5, 0.0, 4.5, 6.0] and numel == 5 * 3 + 3 + 3 * 2 + 2
