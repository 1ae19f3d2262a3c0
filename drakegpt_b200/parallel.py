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


def shard_range(n_live, world, rank, align=64):
    """Contiguous, ``align``-element aligned shard [lo, hi) of a flat arena of ``n_live`` elements owned by ``rank``.
    The shards of ranks 0 .. world-1 tile [0, n_live) exactly (n_live is itself a multiple of ``align``)."""
    if n_live % align != 0:
        raise ValueError(f"arena length {n_live} is not a multiple of {align}")
    units = n_live // align
    base, extra = divmod(units, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo * align, hi * align


class PeerAdamW:
    """Data-parallel optimizer step as one kernel over NVLink peer memory (``dgpt_dp_adamw``): reduce-scatter of the
    gradient arenas + AdamW on this rank's 1/N shard + all-gather of the updated fp32 parameters and bf16 shadows.

    Takes the place of ``GradAllReducer`` + ``FusedAdamW.launch`` in the training step (same ``bucket_ready`` /
    ``finish`` surface, plus ``launch_update``).  Construction exchanges CUDA IPC handles of the gradient, parameter,
    shadow and flag buffers between the ranks of ``group`` (one process per GPU on one NVLink / NVSwitch node).
    The Adam moments of this rank are only meaningful inside its shard; ``gather_moments`` reassembles the full
    vectors (checkpointing).
    """

    fused_optimizer = True

    def __init__(self, flat, opt, group=None, shadow_only=()):
        """``shadow_only``: names of parameters that the training step reads ONLY through their bf16 shadows (the
        tensor-core GEMM weights).  Their updated fp32 masters are not broadcast (2/3 of the outbound NVLink bytes):
        every rank keeps them current for its own shard only and ``sync_master()`` fetches the rest on demand
        (checkpoints, ``state_dict``).  DGPT_DP_BCAST=all broadcasts every fp32 value instead."""
        import ctypes as C
        from . import _lib
        self.flat, self.opt, self.group = flat, opt, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("PeerAdamW supports up to 8 ranks (one NVSwitch node)")
        dev = flat.device
        self.lo, self.hi = shard_range(flat.n_live, self.world, self.rank)
        self.flags = torch.zeros(16, device=dev, dtype=torch.int32)
        self.epoch = torch.zeros(1, device=dev, dtype=torch.int32)
        self.scratch = torch.zeros(2, device=dev, dtype=torch.int32)
        self.need32 = None
        if shadow_only and flat.shadow is not None and os.environ.get("DGPT_DP_BCAST", "shadow") != "all":
            need = torch.ones((flat.n_total + 63) // 64, dtype=torch.uint8)
            for name in shadow_only:
                o, k, _ = flat.slots[name]
                need[o // 64:(o + k + 63) // 64] = 0  # (slots are 64-element aligned: a block belongs to one tensor)
            self.need32 = need.to(dev)
        self.master_stale = False
        lib = _lib.lib()
        mine = []
        self._local = [flat.g, flat.p, flat.shadow, self.flags]
        for t in self._local:
            if t is None:
                mine.append(None)
                continue
            h = (C.c_ubyte * 64)()
            off = C.c_int64(0)
            _lib.check(lib.dgpt_ipc_export(t.data_ptr(), h, C.byref(off)), "dgpt_ipc_export")
            mine.append((bytes(h), int(off.value)))
        torch.cuda.synchronize(dev)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._opened = []
        self._peer_p = [None] * self.world
        ptrs = (C.c_void_p * (4 * self.world + 1))()
        for r in range(self.world):
            for k in range(4):
                if r == self.rank:
                    ptrs[k * self.world + r] = None if self._local[k] is None else self._local[k].data_ptr()
                elif everyone[r][k] is None:
                    ptrs[k * self.world + r] = None
                else:
                    handle, off = everyone[r][k]
                    out = C.c_void_p()
                    _lib.check(lib.dgpt_ipc_open(handle, off, C.byref(out)), "dgpt_ipc_open")
                    self._opened.append((out.value, off))
                    ptrs[k * self.world + r] = out.value
                    if k == 1:
                        self._peer_p[r] = out.value
        ptrs[4 * self.world] = None if self.need32 is None else self.need32.data_ptr()
        self._ptrs = ptrs
        opt.grad_scale = 1.0 / self.world
        opt.upload()
        flat.before_refresh = self.sync_master  # a shadow re-cast must not read stale fp32 masters
        dist.barrier(group=group)  # every rank has mapped every buffer before any kernel touches them

    @property
    def grad_scale(self):
        return 1.0 / self.world

    def bucket_ready(self):
        pass

    def finish(self):
        pass

    def launch_update(self):
        """Enqueue the fused reduce-scatter + AdamW + all-gather, then clear the local gradient arena."""
        from . import _lib, ops
        f, o = self.flat, self.opt
        _lib.check(_lib.lib().dgpt_dp_adamw(self._ptrs, self.world, self.rank, f.m[self.lo:].data_ptr(),
                                            f.v[self.lo:].data_ptr(), self.lo, self.hi, o.hyper.data_ptr(),
                                            o.step_dev.data_ptr(), self.epoch.data_ptr(), self.scratch.data_ptr(), 0,
                                            ops._stream()), "dgpt_dp_adamw")
        f.g[:f.n_live].zero_()
        if self.need32 is not None:
            self.master_stale = True

    def sync_master(self):
        """Fetch the fp32 masters this rank does not own from their owners' arenas (one-sided peer copies over NVLink;
        no collective: any rank may call it alone, e.g. rank 0 before ``state_dict()``).  The caller makes sure no rank
        is inside a training step (a ``dist.barrier()`` after the last step)."""
        if not self.master_stale:
            return
        from . import _lib
        stream = torch.cuda.current_stream(self.flat.device).cuda_stream
        for r in range(self.world):
            if r == self.rank:
                continue
            lo, hi = shard_range(self.flat.n_live, self.world, r)
            _lib.check(_lib.lib().dgpt_peer_copy(self.flat.p.data_ptr() + 4 * lo, self._peer_p[r] + 4 * lo, 4 * (hi - lo), stream),
                       "dgpt_peer_copy")
        torch.cuda.current_stream(self.flat.device).synchronize()
        self.master_stale = False

    def status(self):
        """0 = ok; 1 / 2 = a peer never reached the first / second barrier of some step (host sync)."""
        return int(self.scratch[1].item())

    def gather_moments(self):
        """Full-length (m, v) assembled from every rank's shard (CPU tensors on every rank)."""
        n = self.flat.n_live
        out = []
        for t in (self.flat.m, self.flat.v):
            full = t[:n].clone()
            parts = [None] * self.world
            dist.all_gather_object(parts, (self.lo, self.hi, t[self.lo:self.hi].cpu()), group=self.group)
            full = full.cpu()
            for lo, hi, x in parts:
                full[lo:hi] = x
            out.append(full)
        return out

    def close(self):
        from . import _lib
        for ptr, off in self._opened:
            _lib.lib().dgpt_ipc_close(ptr, off)
        self._opened = []
