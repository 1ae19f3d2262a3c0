"""Data-parallel host logic on CPU: world_size 2, gloo.  Checks that the bucketed, asynchronous
gradient all-reduce used by the training step sums every element of the flat arena exactly once,
that the 1/world mean goes into the optimizer scale (not the gradients), and the env-driven init."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from drakegpt_b200.parallel import GradAllReducer, bucket_ranges, init_from_env
    r, w, _ = init_from_env("gloo")
    assert (r, w) == (rank, world) and dist.get_world_size() == world
    # a miniature arena with the scaled model's bucket structure
    slots, off = {}, 0
    names = ["token_embedding_table.weight", "position_embedding_table.weight"]
    names += [f"blocks.{i}.{n}" for i in range(3) for n in ("sa_head.qkv", "sa_head.proj.weight", "ffwd.net.0.weight", "ln1.weight")]
    names += ["lm_head.weight", "lm_head.bias"]
    g = torch.Generator().manual_seed(0)
    for n in names:
        k = int(torch.randint(10, 500, (1,), generator=g))
        slots[n] = (off, k, (k,))
        off += (k + 63) // 64 * 64
    n_live = off
    groups = [("lm_head.",)] + [(f"blocks.{i}.",) for i in (2, 1, 0)] + [("token_embedding_table.", "position_embedding_table.")]
    ranges = bucket_ranges(slots, n_live, groups)
    grad = torch.arange(n_live, dtype=torch.float32) * (rank + 1)  # rank-dependent gradient
    want = torch.arange(n_live, dtype=torch.float32) * sum(range(1, world + 1))
    for overlap in (False, True):  # one message at the end of backward (default) / bucketed and overlapped
        grad.copy_(torch.arange(n_live, dtype=torch.float32) * (rank + 1))
        red = GradAllReducer(grad, ranges, overlap=overlap)
        assert red.grad_scale == 1.0 / world and red.overlap == overlap
        for step in range(2):  # reusable across steps
            if step:
                grad.copy_(torch.arange(n_live, dtype=torch.float32) * (rank + 1))
            for _ in ranges:
                red.bucket_ready()
            red.finish()
            assert torch.equal(grad, want), (rank, step, overlap)
    torch.save(grad, os.path.join(out_dir, f"rank{rank}.pt"))
    red.bucket_ready()  # (leaves one asynchronous bucket in flight: drained below before the group goes away)
    try:
        red.finish()
        raise AssertionError("finish() must refuse a partial reduction")
    except RuntimeError:
        pass
    for w in red._pending:
        w.wait()
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a, b = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert torch.equal(a, b)  # replicas hold identical reduced gradients


def test_single_process_reducer_is_a_noop():
    from drakegpt_b200.parallel import GradAllReducer
    g = torch.ones(128)
    red = GradAllReducer(g, [(0, 64), (64, 128)])
    assert red.world == 1 and red.grad_scale == 1.0
    red.bucket_ready()
    red.bucket_ready()
    red.finish()
    assert torch.equal(g, torch.ones(128))
    with pytest.raises(IndexError):
        red.bucket_ready(); red.bucket_ready(); red.bucket_ready()
