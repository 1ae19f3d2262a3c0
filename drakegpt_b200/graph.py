"""CUDA-graph capture of the whole training step (forward + backward + all-reduce + AdamW).

Python/ctypes launch overhead (about 150 kernel launches per step) would dominate a ~2 ms step,
so the step is recorded once and replayed.  Everything that changes per step lives in device
memory: the token batch (static input buffers), the AdamW step counter and the dropout seed
offset (both bumped by tiny kernels inside the graph).
"""
import torch


class GraphedTrainStep:
    def __init__(self, runner, batch_size, seq_len, reducer=None, warmup=2):
        if runner.opt is None:
            raise RuntimeError("configure_optimizer() before capturing the step")
        self.runner = runner
        self.reducer = reducer
        dev = runner.device
        self.idx = torch.zeros((batch_size, seq_len), device=dev, dtype=torch.int64)
        self.targets = torch.zeros((batch_size, seq_len), device=dev, dtype=torch.int64)
        flat, opt = runner.flat, runner.opt
        self.reducer = reducer
        self._lazy_master = reducer is not None and getattr(reducer, "need32", None) is not None
        opt.upload()
        keep = [t.clone() for t in (flat.p, flat.m, flat.v, opt.step_dev, runner.seed_dev)]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):  # allocate every workspace buffer, load kernels, init NCCL channels
                runner.train_step_launch(self.idx, self.targets, reducer)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = runner.train_step_launch(self.idx, self.targets, reducer)
        # capture records without executing; undo the warm-up steps
        for dst, src in zip((flat.p, flat.m, flat.v, opt.step_dev, runner.seed_dev), keep):
            dst.copy_(src)
        flat.g.zero_()
        if reducer is not None and getattr(reducer, "fused_optimizer", False):
            reducer.master_stale = False  # every rank restored its own complete copy: nothing to fetch from peers
        flat.refresh_shadow(force=True)  # bf16 shadows follow the restored fp32 masters
        torch.cuda.synchronize(dev)

    def step(self, idx=None, targets=None):
        """Replay one training step.  ``idx``/``targets`` may be device or pinned-host tensors."""
        if idx is not None:
            self.idx.copy_(idx, non_blocking=True)
            self.targets.copy_(targets, non_blocking=True)
        self.graph.replay()
        self.runner.opt.t += 1
        if self._lazy_master:  # the replayed fused optimizer step left the peers' fp32 masters behind (see PeerAdamW.sync_master)
            self.reducer.master_stale = True
        return self.loss


class GraphedEvalStep:
    """CUDA-graph capture of the evaluation forward (reference: src/train.py:61-75).

    One replay = eval-mode forward of a (B, T) batch + ``loss_sum += loss`` on the device, so a whole split of
    ``evaluate_loss`` costs one graph replay per batch and ONE host sync, instead of the reference's
    ``loss.item()`` per batch.  The batch lives in static device buffers filled by ``copy_`` before each replay.
    """

    def __init__(self, runner, batch_size, seq_len):
        self.runner = runner
        dev = runner.device
        self.idx = torch.zeros((batch_size, seq_len), device=dev, dtype=torch.int64)
        self.targets = torch.zeros((batch_size, seq_len), device=dev, dtype=torch.int64)
        self.loss_sum = torch.zeros((), device=dev, dtype=torch.float32)
        self.flat_gen = runner._flat_gen
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            runner.forward(self.idx, self.targets, training=False, want_logits=False)  # allocate workspaces, load kernels
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            _, loss = runner.forward(self.idx, self.targets, training=False, want_logits=False)
            self.loss_sum.add_(loss)
        self.loss_sum.zero_()

    def reset(self):
        self.loss_sum.zero_()

    def step(self, idx, targets):
        self.idx.copy_(idx, non_blocking=True)
        self.targets.copy_(targets, non_blocking=True)
        self.runner.flat.refresh_shadow()  # parameters touched by torch since the capture (e.g. load_state_dict)
        self.graph.replay()
