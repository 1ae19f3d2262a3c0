"""Data parallelism for the training step: one process per GPU, NCCL over NVLink/NVSwitch.

The step shards over the batch: every rank holds a full replica, draws its own
batch, and the only exchange is the SUM of the flat fp32 gradient arena.  ``GradAllReducer``
issues that all-reduce in buckets, each launched asynchronously as soon as the backward pass
has finished the corresponding slice of the arena (lm_head first, then blocks L-1 .. 0, then the
embeddings), so the transfers overlap the rest of backward; the 1/world_size mean is folded into
the fused AdamW (``grad_scale``), not applied to the gradients.

Works with any ``torch.distributed`` backend: ``nccl`` on the GPUs, ``gloo`` in the CPU tests.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* variables."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def bucket_ranges(slots, n_live, groups):
    """Arena ranges [lo, hi) per bucket; ``groups`` is a list of name-prefix tuples in completion order.

    Every live parameter must fall in exactly one group; ranges are contiguous because the arena
    follows registration order.
    """
    out = []
    seen = 0
    for prefixes in groups:
        names = [n for n in slots if n.startswith(tuple(prefixes)) and slots[n][0] < n_live]
        if not names:
            continue
        lo = min(slots[n][0] for n in names)
        hi = max(slots[n][0] + ((slots[n][1] + 63) // 64) * 64 for n in names)
        out.append((lo, min(hi, n_live)))
        seen += len(names)
    live = [n for n in slots if slots[n][0] < n_live]
    if seen != len(live):
        raise ValueError("bucket groups do not cover every trainable parameter exactly once")
    return out


class GradAllReducer:
    """Bucketed, asynchronous SUM all-reduce of a flat gradient buffer."""

    def __init__(self, grad_buffer, ranges, group=None):
        self.g = grad_buffer
        self.ranges = list(ranges)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._pending = []
        self._next = 0

    @property
    def grad_scale(self):
        return 1.0 / self.world

    def bucket_ready(self):
        """Called by the backward pass each time the next bucket's gradients are final."""
        lo, hi = self.ranges[self._next]
        self._next += 1
        if self.world > 1:
            self._pending.append(dist.all_reduce(self.g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Make the current stream wait for every outstanding bucket (before the optimizer)."""
        if self._next != len(self.ranges):
            raise RuntimeError(f"only {self._next} of {len(self.ranges)} gradient buckets were reduced")
        for w in self._pending:
            w.wait()
        self._pending = []
        self._next = 0
