"""CPU oracle for the DrakeGPT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, torch-CPU fp32 restatement of the reference's training step and
generation (ChrisTho23/DrakeGPT).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
module, and only as the checker / the timed CPU baseline; nothing under
``drakegpt_b200/`` imports it.

The arithmetic of the reference lives in PyTorch (pinned torch 2.2.1,
poetry.lock:1159); the reference itself only composes ``nn.Linear``,
``nn.Embedding``, ``nn.LayerNorm``, ``F.softmax``, ``nn.Dropout``,
``F.cross_entropy``, ``torch.multinomial``, ``AdamW`` and ``CyclicLR``.  This
file restates that composition as plain functions over a reference-layout
``state_dict`` (same keys/shapes as ``model/*.pt``), looping over heads one at
a time exactly like the reference does, so it is also a fair "port" CPU
baseline.

Parity pin: ``tests/golden/*.pt`` hold outputs of the *real* reference code
(imported from /root/reference/src by ``tests/golden/make_golden.py`` in the
build container) on the six shipped checkpoints and on seeded inits;
``tests/test_oracle_golden.py`` checks this oracle against every one of them.
The reference ships no tests of its own, so these
reference-generated vectors are the pin.

Citations are file:line under the reference repository.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

KINDS = (
    "BigramLM",
    "SingleHeadAttentionLM",
    "MultiHeadAttentionLM",
    "BlocksLM",
    "ResidualBlocksLM",
    "TransformerLM",
)


# --------------------------------------------------------------------------- #
# tokenizer (src/preprocessing.py:3-26)
# --------------------------------------------------------------------------- #
def get_mapper(text):
    """vocab = sorted set of characters; encode str->list[int]; decode back.

    Follows src/preprocessing.py:14-26.
    """
    chars = sorted(set(text))
    to_id = {}
    for i, ch in enumerate(chars):
        to_id[ch] = i

    def encode(s):
        return [to_id[c] for c in s]

    def decode(ids):
        return "".join(chars[i] for i in ids)

    return encode, decode, len(chars)


def get_batch(data, context_length, batch_size, generator=None):
    """Random windows + next-token targets (src/preprocessing.py:43-45)."""
    ix = torch.randint(len(data) - context_length, (batch_size,), generator=generator)
    x = torch.stack([data[i : i + context_length] for i in ix])
    y = torch.stack([data[i + 1 : i + context_length + 1] for i in ix])
    return x, y


# --------------------------------------------------------------------------- #
# model structure helpers
# --------------------------------------------------------------------------- #
def state_dict_schema(kind, vocab_size, embedding_dim=32, context_length=8,
                      head_size=32, num_heads=4, num_layers=3):
    """Key -> shape, in the reference's registration order.

    Mirrors the constructors at src/model.py:75-78,152-167,251-271,355-372,
    456-473,558-576 and src/model_component.py:24-34,83-85,116-121,155-164,
    195-201,236-239,280-289,318-325,365-376,428-434,477-489.
    """
    V, C, T = vocab_size, embedding_dim, context_length
    out = OrderedDict()
    if kind == "BigramLM":
        out["token_embedding_table.weight"] = (V, V)
        return out
    out["token_embedding_table.weight"] = (V, C)
    out["position_embedding_table.weight"] = (T, C)

    def head(pfx, H):
        out[pfx + "tril"] = (T, T)
        out[pfx + "key.weight"] = (H, C)
        out[pfx + "query.weight"] = (H, C)
        out[pfx + "value.weight"] = (H, C)

    if kind == "SingleHeadAttentionLM":
        head("sa_head.", head_size)
    elif kind == "MultiHeadAttentionLM":
        for j in range(num_heads):
            head(f"sa_head.heads.{j}.", head_size // num_heads)  # src/model.py:262-267
    else:
        H = C // num_heads
        for i in range(num_layers):
            b = f"blocks.{i}."
            for j in range(num_heads):
                head(f"{b}sa_head.heads.{j}.", H)
            if kind in ("ResidualBlocksLM", "TransformerLM"):
                out[b + "sa_head.proj.weight"] = (C, C)
                out[b + "sa_head.proj.bias"] = (C,)
            if kind == "BlocksLM":
                out[b + "ffwd.net.0.weight"] = (C, C)
                out[b + "ffwd.net.0.bias"] = (C,)
            else:
                out[b + "ffwd.net.0.weight"] = (4 * C, C)
                out[b + "ffwd.net.0.bias"] = (4 * C,)
                out[b + "ffwd.net.2.weight"] = (C, 4 * C)
                out[b + "ffwd.net.2.bias"] = (C,)
            if kind == "TransformerLM":
                for n in ("ln1", "ln2"):
                    out[b + n + ".weight"] = (C,)
                    out[b + n + ".bias"] = (C,)
        if kind == "TransformerLM":
            out["ln_f.weight"] = (C,)
            out["ln_f.bias"] = (C,)
    out["lm_head.weight"] = (V, C)
    out["lm_head.bias"] = (V,)
    return out


def synthetic_state_dict(kind, seed, **cfg):
    """Deterministic weights from a formula both sides can regenerate.

    Same distributions as the reference's constructors draw (nn.Linear: weight and
    bias ~ U(+-1/sqrt(fan_in)); nn.Embedding ~ N(0,1); nn.LayerNorm 1/0, here with a
    small perturbation so the affine part is exercised), but generated key by key
    from one seeded generator instead of in torch-constructor order.  Used for
    shapes whose weights cannot be shipped as fixtures (TransformerLM_scaled).
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    T = cfg.get("context_length", 8)
    schema = state_dict_schema(kind, **cfg)
    for key, shape in schema.items():
        if key.endswith("tril"):
            sd[key] = torch.tril(torch.ones(T, T))
        elif ".ln" in key or key.startswith("ln_f"):
            base = 1.0 if key.endswith("weight") else 0.0
            sd[key] = base + 0.05 * torch.randn(shape, generator=g)
        elif "embedding" in key:
            sd[key] = torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] if len(shape) == 2 else schema[key[:-4] + "weight"][1]
            bound = 1.0 / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
    return sd


def infer_config(kind, sd):
    """Recover constructor arguments from a reference-layout state_dict."""
    cfg = {"vocab_size": sd["token_embedding_table.weight"].shape[0]}
    if kind == "BigramLM":
        return cfg
    cfg["embedding_dim"] = sd["token_embedding_table.weight"].shape[1]
    cfg["context_length"] = sd["position_embedding_table.weight"].shape[0]
    if kind == "SingleHeadAttentionLM":
        cfg["head_size"] = sd["sa_head.key.weight"].shape[0]
        return cfg
    if kind == "MultiHeadAttentionLM":
        nh = len([k for k in sd if k.endswith("key.weight")])
        cfg["num_heads"] = nh
        cfg["head_size"] = nh * sd["sa_head.heads.0.key.weight"].shape[0]
        return cfg
    nh = len([k for k in sd if k.startswith("blocks.0.") and k.endswith("key.weight")])
    nl = len([k for k in sd if k.endswith("heads.0.key.weight")])
    cfg["num_heads"], cfg["num_layers"] = nh, nl
    return cfg


# --------------------------------------------------------------------------- #
# components
# --------------------------------------------------------------------------- #
def _one_head(sd, pfx, x, p_drop, training):
    """Single causal self-attention head.

    src/model_component.py:36-66 (Head) and :378-407 (Head2: + dropout on the
    probabilities at :401).  scale = head_size ** -0.5 applied after q @ k^T
    (:27,:56); mask = tril[:T,:T] == 0 -> -inf (:57-59); softmax over the last
    dim (:60); value projection and weighted sum (:63-64).
    """
    T = x.shape[1]
    wk, wq, wv = sd[pfx + "key.weight"], sd[pfx + "query.weight"], sd[pfx + "value.weight"]
    k = x @ wk.t()
    q = x @ wq.t()
    scores = (q @ k.transpose(-2, -1)) * (wk.shape[0] ** -0.5)
    scores = scores.masked_fill(sd[pfx + "tril"][:T, :T] == 0, float("-inf"))
    probs = F.softmax(scores, dim=-1)
    if p_drop is not None:
        probs = F.dropout(probs, p_drop, training)
    v = x @ wv.t()
    return probs @ v


def _heads(sd, pfx, x, p_drop, training):
    """Python loop over heads + concat (src/model_component.py:103,260,453)."""
    n = len([k for k in sd if k.startswith(pfx + "heads.") and k.endswith("key.weight")])
    outs = [_one_head(sd, f"{pfx}heads.{j}.", x, p_drop, training) for j in range(n)]
    return torch.cat(outs, dim=-1)


def _linear(sd, pfx, x):
    return x @ sd[pfx + "weight"].t() + sd[pfx + "bias"]


def _layer_norm(sd, pfx, x):
    """nn.LayerNorm(C): eps 1e-5, biased variance, affine."""
    return F.layer_norm(x, (x.shape[-1],), sd[pfx + "weight"], sd[pfx + "bias"], 1e-5)


def _block(kind, sd, b, x, p_drop, training):
    if kind == "BlocksLM":
        # src/model_component.py:179-181: ffwd(sa_head(x)), no residual;
        # FeedForward = Linear(C,C)+ReLU (:118-121)
        a = _heads(sd, b + "sa_head.", x, None, training)
        return torch.relu(_linear(sd, b + "ffwd.net.0.", a))
    if kind == "ResidualBlocksLM":
        # src/model_component.py:304-306; MultiHeadAttention2 (:259-261);
        # FeedForward2 (:197-201)
        a = _linear(sd, b + "sa_head.proj.", _heads(sd, b + "sa_head.", x, None, training))
        x = x + a
        h = torch.relu(_linear(sd, b + "ffwd.net.0.", x))
        return x + _linear(sd, b + "ffwd.net.2.", h)
    # TransformerLM / ResidualBlock2: src/model_component.py:505-506, pre-LN;
    # MultiHeadAttention3 (:453-454) dropout(proj(cat)); FeedForward3 (:320-325)
    a = _heads(sd, b + "sa_head.", _layer_norm(sd, b + "ln1.", x), p_drop, training)
    a = F.dropout(_linear(sd, b + "sa_head.proj.", a), p_drop, training)
    x = x + a
    h = torch.relu(_linear(sd, b + "ffwd.net.0.", _layer_norm(sd, b + "ln2.", x)))
    return x + F.dropout(_linear(sd, b + "ffwd.net.2.", h), p_drop, training)


def forward(kind, sd, idx, targets=None, dropout=0.0, training=False):
    """``LM.forward(idx, targets) -> (logits, loss)``.

    src/model.py:80-105 (Bigram), :169-200, :273-304, :374-405, :475-506,
    :578-609 (TransformerLM; ln_f constructed at :572 but never applied).
    With targets the returned logits are (B*T, V) (src/model.py:604-606).
    """
    tok = sd["token_embedding_table.weight"]
    if kind == "BigramLM":
        logits = tok[idx]
    else:
        B, T = idx.shape
        x = tok[idx] + sd["position_embedding_table.weight"][torch.arange(T)]
        if kind == "SingleHeadAttentionLM":
            x = _one_head(sd, "sa_head.", x, None, training)
        elif kind == "MultiHeadAttentionLM":
            x = _heads(sd, "sa_head.", x, None, training)
        else:
            nl = len([k for k in sd if k.endswith("heads.0.key.weight")])
            p = dropout if kind == "TransformerLM" else None
            for i in range(nl):
                x = _block(kind, sd, f"blocks.{i}.", x, p, training)
        logits = _linear(sd, "lm_head.", x)
    if targets is None:
        return logits, None
    B, T, V = logits.shape
    logits = logits.view(B * T, V)
    loss = F.cross_entropy(logits, targets.view(B * T))
    return logits, loss


def context_length_of(kind, sd):
    return None if kind == "BigramLM" else sd["position_embedding_table.weight"].shape[0]


@torch.no_grad()
def generate(kind, sd, idx, max_new_tokens, greedy=False, generator=None):
    """Crop to the last context_length tokens, full forward, sample the last step.

    src/model.py:611-636 (and :107-130 for BigramLM, which does not crop).
    ``greedy=True`` replaces softmax->multinomial with argmax (test-harness
    mode; the reference always samples, src/model.py:630-632).
    """
    ctx = context_length_of(kind, sd)
    for _ in range(max_new_tokens):
        cond = idx if ctx is None else idx[:, -ctx:]
        logits, _ = forward(kind, sd, cond)
        last = logits[:, -1, :]
        if greedy:
            nxt = last.argmax(dim=-1, keepdim=True)
        else:
            nxt = torch.multinomial(F.softmax(last, dim=-1), 1, generator=generator)
        idx = torch.cat((idx, nxt), dim=1)
    return idx


# --------------------------------------------------------------------------- #
# optimizer (src/train.py:121-126,149-151; torch.optim.AdamW defaults)
# --------------------------------------------------------------------------- #
def trainable_keys(sd):
    return [k for k in sd if not k.endswith("tril")]


class AdamW:
    """Decoupled-decay Adam exactly as torch.optim.AdamW(lr, betas) computes it.

    One group over every parameter, eps 1e-8, weight_decay 0.01, amsgrad off
    (src/train.py:121).  Parameters whose grad is None (ln_f, Q1)
    are skipped entirely, including their weight decay.
    """

    def __init__(self, sd, lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-2):
        self.sd, self.lr, self.betas, self.eps, self.wd = sd, lr, betas, eps, weight_decay
        self.m = {k: torch.zeros_like(sd[k]) for k in trainable_keys(sd)}
        self.v = {k: torch.zeros_like(sd[k]) for k in trainable_keys(sd)}
        self.t = {k: 0 for k in trainable_keys(sd)}

    @torch.no_grad()
    def step(self, grads):
        b1, b2 = self.betas
        for k, g in grads.items():
            if g is None:
                continue
            p = self.sd[k]
            self.t[k] += 1
            t = self.t[k]
            p.mul_(1.0 - self.lr * self.wd)
            self.m[k].lerp_(g, 1.0 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1.0 - b2)
            bc1 = 1.0 - b1 ** t
            bc2 = 1.0 - b2 ** t
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / bc1)


def loss_and_grads(kind, sd, idx, targets, dropout=0.0, training=True):
    keys = trainable_keys(sd)
    leaves = OrderedDict(sd)
    for k in keys:
        leaves[k] = sd[k].detach().requires_grad_(True)
    logits, loss = forward(kind, leaves, idx, targets, dropout, training)
    gs = torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)
    return logits.detach(), loss.detach(), OrderedDict(zip(keys, gs))


def train_steps(kind, sd, batches, lr, betas=(0.9, 0.95), dropout=0.0, training=True):
    """fwd -> zero_grad -> backward -> AdamW.step per batch (src/train.py:143-151).

    Constant lr: CyclicLR only steps inside the eval branch every 500 iters
    (src/train.py:154-162).  Mutates ``sd`` in place; returns the
    per-step losses (pre-update, as the reference logs them).
    """
    opt = AdamW(sd, lr, betas)
    losses = []
    for x, y in batches:
        _, loss, grads = loss_and_grads(kind, sd, x, y, dropout, training)
        opt.step(grads)
        losses.append(float(loss))
    return losses


def structured_batches(seed, steps, B, T, V=80, corpus_len=65536):
    """Seeded synthetic corpus with learnable structure (every odd token is a function of its predecessor) and
    ``steps`` random (x, y) windows of it, y = x shifted by one like src/preprocessing.py:43-45.  Shared by the
    golden generator (tests/golden/make_golden.py) and the GPU parity tests so both sides see identical batches."""
    g = torch.Generator().manual_seed(seed)
    corpus = torch.randint(0, V, (corpus_len,), generator=g)
    corpus[1::2] = (corpus[::2] * 7 + 3) % V
    out = []
    for _ in range(steps):
        ix = torch.randint(0, corpus_len - T - 1, (B,), generator=g)
        offs = ix.unsqueeze(1) + torch.arange(T + 1).unsqueeze(0)
        win = corpus[offs]
        out.append((win[:, :-1].contiguous(), win[:, 1:].contiguous()))
    return out


def cyclic_lr(step, base_lr, max_lr, step_size_up=5):
    """torch CyclicLR 'triangular' lr after ``step`` scheduler.step() calls.

    src/train.py:122-126 (step_size_up=5, cycle_momentum=False).
    """
    total = 2 * step_size_up
    cycle = math.floor(1 + step / total)
    x = 1.0 + step / total - cycle
    ratio = step_size_up / total
    scale = x / ratio if x <= ratio else (x - 1) / (ratio - 1)
    return base_lr + (max_lr - base_lr) * scale


def model_params(params, model_type, vocab_size):
    """The reference's *estimated* parameter count (src/model.py:8-63)."""
    C, T, L = params["embedding_dim"], params["context_length"], params["num_layers"]
    pos, tok = C * T, C * vocab_size
    kqv, proj = 3 * C * C, C * C
    ffw = 4 * C * C
    block = (C + kqv + proj) + (C + ffw + ffw)
    head = C * vocab_size
    total = tok
    if model_type != "BigramLM":
        total += pos + kqv + head
    if model_type in ("SingleHeadAttentionLM", "MultiHeadAttentionLM"):
        total += pos + head
    if model_type == "BlocksLM":
        total += (kqv + ffw) * L
    if model_type == "ResidualBlocksLM":
        total += (kqv + ffw + ffw) * L
    if model_type == "TransformerLM":
        total += block * L + vocab_size
    return total
