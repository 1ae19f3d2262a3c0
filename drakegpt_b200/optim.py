"""Flat parameter arena + fused AdamW.

``FlatParams`` moves every parameter of a model into ONE contiguous fp32 buffer
(trainable tensors first, never-trained ones such as ``ln_f`` last), with a
matching gradient buffer, Adam moments and -- for the tensor-core path -- a bf16
shadow copy.  ``p.data`` / ``p.grad`` of each ``nn.Parameter`` become views into
those buffers, so torch-side code (``state_dict``, ``torch.optim``) keeps
working, while the fused optimizer, the bf16 re-cast and the data-parallel
gradient all-reduce each touch a single buffer.

``FusedAdamW`` is the reference's ``optimizer.zero_grad(); ...; optimizer.step()``
(src/train.py:121,149,151: AdamW, betas (0.9, 0.95), torch defaults eps 1e-8 and
weight_decay 1e-2 on every parameter) as one kernel launch.
"""
import torch

from . import ops

_ALIGN = 64  # elements; keeps every tensor 256-byte aligned in fp32 and 128-byte aligned in bf16


def _round_up(n, a):
    return (n + a - 1) // a * a


class FlatParams:
    def __init__(self, model, frozen=(), with_shadow=False):
        params = [(n, p) for n, p in model.named_parameters()]
        if not params:
            raise ValueError("model has no parameters")
        dev = params[0][1].device
        if dev.type != "cuda":
            raise ops._lib.KernelError("FlatParams needs the model on a CUDA device (no CPU fallback)")
        frozen = tuple(frozen)
        live = [(n, p) for n, p in params if not n.startswith(frozen)] if frozen else params
        dead = [(n, p) for n, p in params if frozen and n.startswith(frozen)]
        self.slots = {}
        off = 0
        for n, p in live:
            self.slots[n] = (off, p.numel(), tuple(p.shape))
            off += _round_up(p.numel(), _ALIGN)
        self.n_live = off
        for n, p in dead:
            self.slots[n] = (off, p.numel(), tuple(p.shape))
            off += _round_up(p.numel(), _ALIGN)
        self.n_total = off
        self.device = dev
        self.p = torch.zeros(off, device=dev, dtype=torch.float32)
        self.g = torch.zeros(off, device=dev, dtype=torch.float32)
        self.m = torch.zeros(self.n_live, device=dev, dtype=torch.float32)
        self.v = torch.zeros(self.n_live, device=dev, dtype=torch.float32)
        self.shadow = torch.zeros(off, device=dev, dtype=torch.bfloat16) if with_shadow else None
        self.params = dict(params)
        self.frozen_names = [n for n, _ in dead]
        with torch.no_grad():
            for n, p in params:
                o, k, shp = self.slots[n]
                view = self.p[o:o + k].view(shp)
                view.copy_(p.data)
                p.data = view
        self._versions = None
        self.before_refresh = None  # set by parallel.PeerAdamW: fetch fp32 masters owned by other ranks before a re-cast
        self.n_params = sum(k for (_, k, _) in self.slots.values())

    # ---- views ------------------------------------------------------------
    def view(self, name, buf=None):
        o, k, shp = self.slots[name]
        return (self.p if buf is None else buf)[o:o + k].view(shp)

    def grad(self, name):
        return self.view(name, self.g)

    def shadow_of(self, name):
        return self.view(name, self.shadow)

    def is_attached(self):
        """False once something (e.g. ``model.to``) re-pointed a parameter away from the arena."""
        for n, p in self.params.items():
            o, _, _ = self.slots[n]
            if p.data_ptr() != self.p.data_ptr() + 4 * o:
                return False
        return True

    def attach_grads(self):
        """Re-point ``p.grad`` at the arena (after ``zero_grad(set_to_none=True)``)."""
        for n, p in self.params.items():
            if n not in self.frozen_names:
                o, k, shp = self.slots[n]
                p.grad = self.g[o:o + k].view(shp)

    def refresh_shadow(self, force=False):
        """Re-cast the bf16 shadow if any parameter was modified by torch since the last cast."""
        if self.shadow is None:
            return
        vers = tuple(p._version for p in self.params.values())
        if force or vers != self._versions:
            if self.before_refresh is not None:
                self.before_refresh()
            ops.raw_cast_bf16(self.p, self.shadow)
            self._versions = vers


class FusedAdamW:
    """One-kernel AdamW over a ``FlatParams`` arena (decoupled decay, bias correction as torch)."""

    def __init__(self, flat, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-2):
        self.flat = flat
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), tuple(betas), float(eps), float(weight_decay)
        self.t = 0  # host mirror of the device-side step counter
        self.grad_scale = 1.0
        self.hyper = torch.zeros(6, device=flat.device, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=flat.device, dtype=torch.int64)
        self.param_groups = [{"lr": self.lr}]  # enough surface for lr schedulers that poke param_groups
        self._uploaded = None
        self.upload()

    def upload(self):
        """(Re)send lr/betas/eps/wd/grad_scale to the device if they changed (blocking, rare)."""
        self.lr = float(self.param_groups[0]["lr"])
        vals = (self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.grad_scale)
        if vals != self._uploaded:
            self.hyper.copy_(torch.tensor(vals, dtype=torch.float32))
            self._uploaded = vals

    def launch(self, zero_grad=True):
        """Enqueue the update; bias corrections come from the device-side step counter."""
        f = self.flat
        ops.raw_adamw(f.p, f.g, f.m, f.v, f.shadow, self.hyper, self.step_dev, zero_grad=zero_grad, n=f.n_live)

    def step(self, zero_grad=True):
        self.upload()
        self.launch(zero_grad)
        self.t += 1

    def zero_grad(self, set_to_none=False):
        self.flat.g.zero_()
