"""Drop-in counterparts of the reference's ``src/model_component.py`` modules.

Same class names, constructor signatures, ``forward(x:(B,T,C)) -> (B,T,.)``
contract and ``state_dict`` layout (per-head ``key/query/value.weight`` and
``tril`` entries, ``proj``, ``ffwd.net.{0,2}``, ``ln1/ln2``) as the reference,
but every ``forward`` runs hand-written sm_100a kernels through
``drakegpt_b200.ops``: one packed QKV GEMM + one fused causal-attention kernel
for all heads (instead of a Python loop over heads and a ``torch.cat``), and
GEMMs whose epilogues carry bias / ReLU / dropout / residual.

Internally the per-head projection weights live in ONE packed parameter
``qkv`` of shape (3, NH, H, C) ordered (query, key, value); ``state_dict()``
and ``load_state_dict()`` translate to and from the reference's per-head keys
(SURVEY.md Q2, Appendix A).  The causal mask is built into the kernels; the
``tril`` buffers are emitted for checkpoint compatibility and checked on load.

These standalone modules always run the exact fp32 path; the bf16 tensor-core
path is the fused engine behind ``TransformerLM`` (``engine.py``).
"""
import math

import torch
import torch.nn as nn

from . import ops

_QKV_NAMES = ("query", "key", "value")  # packed order
_SD_ORDER = ("key", "query", "value")   # registration order in the reference (model_component.py:28-30)


class _WeightView:
    """``head.key`` / ``head.query`` / ``head.value``: exposes ``.weight`` like the reference's nn.Linear."""

    def __init__(self, core, which, j):
        self._core, self._which, self._j = core, which, j

    @property
    def weight(self):
        return self._core.qkv[self._which, self._j]


class _HeadView:
    """``mha.heads[j]`` -- a window onto head j of the packed attention parameters."""

    def __init__(self, core, j):
        self._core, self._j = core, j
        self.head_size = core.head_size
        self.scale = core.head_size ** -0.5
        self.query, self.key, self.value = (_WeightView(core, i, j) for i in range(3))

    @property
    def tril(self):
        return self._core._tril

    def __call__(self, x):
        c = self._core
        return ops.causal_attention(x, c.qkv[:, self._j: self._j + 1], dropout_p=c.attn_dropout, training=c.training)


class _PackedAttention(nn.Module):
    """Packed multi-head causal self-attention parameters + reference-layout (de)serialisation."""

    def __init__(self, num_heads, head_size, embedding_dim, context_length, attn_dropout=0.0, single=False):
        super().__init__()
        self.num_heads, self.head_size = num_heads, head_size
        self.embedding_dim, self.context_length = embedding_dim, context_length
        self.attn_dropout = float(attn_dropout)
        self._single = single
        self.qkv = nn.Parameter(torch.empty(3, num_heads, head_size, embedding_dim))
        bound = 1.0 / math.sqrt(embedding_dim)  # nn.Linear default: kaiming_uniform(a=sqrt(5))
        nn.init.uniform_(self.qkv, -bound, bound)
        self.register_buffer("_tril", torch.tril(torch.ones(context_length, context_length)), persistent=False)

    # ---- reference-layout state_dict ------------------------------------
    def _head_prefix(self, j):
        return "" if self._single else f"heads.{j}."

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for j in range(self.num_heads):
            hp = prefix + self._head_prefix(j)
            destination[hp + "tril"] = self._tril.detach().clone()
            for name in _SD_ORDER:
                w = self.qkv[_QKV_NAMES.index(name), j]
                destination[hp + name + ".weight"] = w if keep_vars else w.detach().clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        expected = set()
        for j in range(self.num_heads):
            hp = prefix + self._head_prefix(j)
            key = hp + "tril"
            expected.add(key)
            if key in state_dict:
                t = state_dict[key]
                ref = torch.tril(torch.ones(self.context_length, self.context_length))
                if tuple(t.shape) != tuple(ref.shape) or not torch.equal(t.detach().cpu().float(), ref):
                    error_msgs.append(f"{key}: kernels implement the causal (lower-triangular) mask only; "
                                      f"got a different tril buffer of shape {tuple(t.shape)}")
            elif strict:
                missing_keys.append(key)
            for name in _SD_ORDER:
                key = hp + name + ".weight"
                expected.add(key)
                if key not in state_dict:
                    if strict:
                        missing_keys.append(key)
                    continue
                src = state_dict[key]
                which = _QKV_NAMES.index(name)
                if tuple(src.shape) != tuple(self.qkv.shape[2:]):
                    error_msgs.append(f"size mismatch for {key}: copying a param with shape {tuple(src.shape)} "
                                      f"from checkpoint, the shape in current model is {tuple(self.qkv.shape[2:])}.")
                    continue
                with torch.no_grad():
                    # through the Parameter (not .data): bumps qkv._version, which is what tells
                    # optim.FlatParams.refresh_shadow that the bf16 shadow of this tensor is stale
                    self.qkv[which, j].copy_(src)
        if strict:
            children = tuple(prefix + n + "." for n in self._modules)
            for key in state_dict:
                if key.startswith(prefix) and key not in expected and not key.startswith(children):
                    unexpected_keys.append(key)

    def attend(self, x):
        return ops.causal_attention(x, self.qkv, dropout_p=self.attn_dropout, training=self.training)


class Head(_PackedAttention):
    """One causal self-attention head (reference: src/model_component.py:5-66).

    ``Head(head_size, embedding_dim, context_length)``; forward (B,T,C) -> (B,T,H).
    """

    def __init__(self, head_size, embedding_dim, context_length, _dropout=0.0):
        super().__init__(1, head_size, embedding_dim, context_length, _dropout, single=True)
        self.scale = head_size ** -0.5
        self.query, self.key, self.value = (_WeightView(self, i, 0) for i in range(3))

    @property
    def tril(self):
        return self._tril

    def forward(self, x):
        return self.attend(x)


SingleHeadAttention = Head  # README.md:20 name for the same class


class Head2(Head):
    """Head + dropout on the attention probabilities (src/model_component.py:343-407)."""

    def __init__(self, head_size, embedding_dim, context_length, dropout):
        super().__init__(head_size, embedding_dim, context_length, dropout)
        self.dropout = nn.Dropout(dropout)


class MultiHeadAttention(_PackedAttention):
    """NH heads + concat (src/model_component.py:69-103); forward (B,T,C) -> (B,T,NH*H)."""

    def __init__(self, num_heads, head_size, embedding_dim, context_length, _dropout=0.0):
        super().__init__(num_heads, head_size, embedding_dim, context_length, _dropout)

    @property
    def heads(self):
        return [_HeadView(self, j) for j in range(self.num_heads)]

    def forward(self, x):
        return self.attend(x)


class MultiHeadAttention2(MultiHeadAttention):
    """+ output projection Linear(C,C) (src/model_component.py:220-261)."""

    def __init__(self, num_heads, head_size, embedding_dim, context_length, _dropout=0.0):
        super().__init__(num_heads, head_size, embedding_dim, context_length, _dropout)
        self.proj = nn.Linear(embedding_dim, embedding_dim)

    def forward(self, x, _residual=None):
        return ops.linear(self.attend(x), self.proj.weight, self.proj.bias, residual=_residual)


class MultiHeadAttention3(MultiHeadAttention2):
    """Head2 heads + dropout(proj(.)) (src/model_component.py:409-455)."""

    def __init__(self, num_heads, head_size, embedding_dim, context_length, dropout):
        super().__init__(num_heads, head_size, embedding_dim, context_length, dropout)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, _residual=None):
        return ops.linear(self.attend(x), self.proj.weight, self.proj.bias, residual=_residual,
                          dropout_p=self.dropout.p, training=self.training)


class FeedForward(nn.Module):
    """ReLU(Linear(C,C)) (src/model_component.py:106-137)."""

    def __init__(self, embedding_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(embedding_dim, embedding_dim), nn.ReLU())

    def forward(self, x):
        return ops.linear(x, self.net[0].weight, self.net[0].bias, relu=True)


class FeedForward2(nn.Module):
    """Linear(C,4C) -> ReLU -> Linear(4C,C) (src/model_component.py:184-217)."""

    def __init__(self, embedding_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(embedding_dim, 4 * embedding_dim), nn.ReLU(),
                                 nn.Linear(4 * embedding_dim, embedding_dim))

    def forward(self, x, _residual=None):
        h = ops.linear(x, self.net[0].weight, self.net[0].bias, relu=True)
        return ops.linear(h, self.net[2].weight, self.net[2].bias, residual=_residual)


class FeedForward3(nn.Module):
    """FeedForward2 + Dropout (src/model_component.py:308-340)."""

    def __init__(self, embedding_dim, dropout):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(embedding_dim, 4 * embedding_dim), nn.ReLU(),
                                 nn.Linear(4 * embedding_dim, embedding_dim), nn.Dropout(dropout))

    def forward(self, x, _residual=None):
        h = ops.linear(x, self.net[0].weight, self.net[0].bias, relu=True)
        return ops.linear(h, self.net[2].weight, self.net[2].bias, residual=_residual,
                          dropout_p=self.net[3].p, training=self.training)


class Block(nn.Module):
    """ffwd(sa_head(x)), no residual (src/model_component.py:140-181)."""

    def __init__(self, embedding_dim, context_length, num_heads):
        super().__init__()
        self.sa_head = MultiHeadAttention(num_heads, embedding_dim // num_heads, embedding_dim, context_length)
        self.ffwd = FeedForward(embedding_dim)

    def forward(self, x):
        return self.ffwd(self.sa_head(x))


class ResidualBlock(nn.Module):
    """x += MHA2(x); x += FF2(x) (src/model_component.py:264-306).  Note the argument order."""

    def __init__(self, embedding_dim, num_heads, context_length):
        super().__init__()
        self.sa_head = MultiHeadAttention2(num_heads, embedding_dim // num_heads, embedding_dim, context_length)
        self.ffwd = FeedForward2(embedding_dim)

    def forward(self, x):
        x = self.sa_head(x, _residual=x)  # residual add fused into the projection GEMM epilogue
        return self.ffwd(x, _residual=x)


class ResidualBlock2(nn.Module):
    """Pre-LN transformer block with dropout (src/model_component.py:458-507)."""

    def __init__(self, embedding_dim, num_heads, context_length, dropout):
        super().__init__()
        self.sa_head = MultiHeadAttention3(num_heads, embedding_dim // num_heads, embedding_dim, context_length,
                                           dropout)
        self.ffwd = FeedForward3(embedding_dim, dropout)
        self.ln1 = nn.LayerNorm(embedding_dim)
        self.ln2 = nn.LayerNorm(embedding_dim)

    def forward(self, x):
        x = self.sa_head(ops.layer_norm(x, self.ln1.weight, self.ln1.bias, self.ln1.eps), _residual=x)
        return self.ffwd(ops.layer_norm(x, self.ln2.weight, self.ln2.bias, self.ln2.eps), _residual=x)
