"""Data parallelism for the training step: one process per GPU, NCCL over NVLink/NVSwitch.

The step shards over the batch: every rank holds a full replica, draws its own
batch, and the only exchange is the SUM of the flat fp32 gradient arena.  ``GradAllReducer``
issues it either as ONE message when backward has finished (default: fastest on NVSwitch, see the
class) or in buckets launched asynchronously as soon as the backward pass has finished the
corresponding slice of the arena (lm_head first, then blocks L-1 .. 0, then the embeddings);
the 1/world_size mean is folded into the fused AdamW (``grad_scale``), not applied to the gradients.

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
    """SUM all-reduce of a flat gradient buffer, either overlapped with backward or as one message at its end.

    ``overlap=True``: one asynchronous all-reduce per bucket, launched the moment the backward pass has finished
    that slice (lm_head first, then blocks L-1 .. 0, then the embeddings).
    ``overlap=False`` (default): ``bucket_ready`` only counts; ``finish`` reduces the whole live arena in ONE call.
    Measured on 8 x B200 (NVLink 5 / NVSwitch, 43 MB of fp32 gradients, 2.4 ms compute step): the single message
    costs ~0.2 ms, whereas the overlapped buckets cost 0.5 - 1.0 ms -- every kernel of the step is a persistent
    grid sized to all 148 SMs, so NCCL kernels running beside them either wait for SMs or push a GEMM's last
    CTAs into a second wave.
    """

    def __init__(self, grad_buffer, ranges, group=None, overlap=None):
        self.g = grad_buffer
        self.ranges = list(ranges)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if overlap is None:
            overlap = os.environ.get("DGPT_DP_OVERLAP", "0") == "1"
        self.overlap = bool(overlap)
        self._pending = []
        self._next = 0

    @property
    def grad_scale(self):
        return 1.0 / self.world

    def bucket_ready(self):
        """Called by the backward pass each time the next bucket's gradients are final."""
        lo, hi = self.ranges[self._next]
        self._next += 1
        if self.world > 1 and self.overlap:
            self._pending.append(dist.all_reduce(self.g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Make the current stream wait for the reduced gradients (before the optimizer)."""
        if self._next != len(self.ranges):
            raise RuntimeError(f"only {self._next} of {len(self.ranges)} gradient buckets were reduced")
        if self.world > 1 and not self.overlap and self.ranges:
            lo, hi = min(r[0] for r in self.ranges), max(r[1] for r in self.ranges)
            dist.all_reduce(self.g[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        for w in self._pending:
            w.wait()
        self._pending = []
        self._next = 0
