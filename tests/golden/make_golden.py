"""Generate golden vectors from the REAL reference (ChrisTho23/DrakeGPT).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports the reference's own ``src/model.py`` / ``src/model_component.py`` /
``src/preprocessing.py`` unmodified, runs them on CPU fp32 and stores small
input/output fixtures under ``tests/golden/``.  The GPU box has no
/root/reference; tests there read only these fixtures.
"""
import os
import sys
from collections import OrderedDict

import torch

REF = os.environ.get("DRAKE_REF", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src"))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import model as ref_model  # noqa: E402
import model_component as ref_mc  # noqa: E402
import preprocessing as ref_pre  # noqa: E402
from oracle import drake_oracle as O  # noqa: E402

P = dict(context_length=8, batch_size=32, embedding_dim=32, head_size=32, num_heads=4,
         num_layers=3, dropout=0.1)
V = 80


def build(kind, params=P, vocab=V):
    C, T = params["embedding_dim"], params["context_length"]
    if kind == "BigramLM":
        return ref_model.BigramLM(vocab)
    if kind == "SingleHeadAttentionLM":
        return ref_model.SingleHeadAttentionLM(vocab, C, T, params["head_size"])
    if kind == "MultiHeadAttentionLM":
        return ref_model.MultiHeadAttentionLM(vocab, C, T, params["head_size"], params["num_heads"])
    if kind == "BlocksLM":
        return ref_model.BlocksLM(vocab, C, T, params["num_heads"], params["num_layers"])
    if kind == "ResidualBlocksLM":
        return ref_model.ResidualBlocksLM(vocab, C, T, params["num_heads"], params["num_layers"])
    return ref_model.TransformerLM(vocab, C, T, params["num_heads"], params["num_layers"], params["dropout"])


def greedy(m, idx, n):
    ctx = getattr(m, "context_length", None)
    with torch.no_grad():
        for _ in range(n):
            cond = idx if ctx is None else idx[:, -ctx:]
            logits, _ = m(cond)
            idx = torch.cat((idx, logits[:, -1, :].argmax(-1, keepdim=True)), 1)
    return idx


def checkpoints():
    os.makedirs(os.path.join(HERE, "checkpoints"), exist_ok=True)
    out = {}
    for kind in O.KINDS:
        sd = torch.load(os.path.join(REF, "model", kind + ".pt"), map_location="cpu", weights_only=True)
        sd = OrderedDict((k, v.detach().clone().contiguous()) for k, v in sd.items())
        torch.save(sd, os.path.join(HERE, "checkpoints", kind + ".pt"))
        m = build(kind)
        m.load_state_dict(sd, strict=True)
        m.eval()
        rec = {"cases": []}
        g = torch.Generator().manual_seed(1234)
        for (B, T) in [(4, 8), (2, 1), (3, 3), (32, 8)]:
            x = torch.randint(0, V, (B, T), generator=g)
            y = torch.randint(0, V, (B, T), generator=g)
            with torch.no_grad():
                lg, _ = m(x)
                lg2, loss = m(x, y)
            assert lg2.shape == (B * T, V)
            rec["cases"].append({"idx": x, "targets": y, "logits": lg.clone(), "loss": loss.clone()})
        rec["greedy_1"] = greedy(m, torch.zeros((1, 1), dtype=torch.long), 64)
        start = torch.tensor([[0], [14], [30]], dtype=torch.long)
        rec["greedy_3"] = greedy(m, start, 40)
        # gradients of the loss on the (32,8) case
        m.train() if kind != "TransformerLM" else m.eval()  # keep dropout off for exact grads
        x, y = rec["cases"][3]["idx"], rec["cases"][3]["targets"]
        m.zero_grad()
        _, loss = m(x, y)
        loss.backward()
        rec["grads"] = {k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}
        out[kind] = rec
    torch.save(out, os.path.join(HERE, "ckpt_vectors.pt"))


def train_curves():
    """200 AdamW steps from the shipped checkpoints on seeded synthetic batches."""
    out = {}
    g = torch.Generator().manual_seed(777)
    batches = [(torch.randint(0, V, (32, 8), generator=g), torch.randint(0, V, (32, 8), generator=g))
               for _ in range(200)]
    out["batches_seed"] = 777
    for kind, p_drop in [("BigramLM", None), ("SingleHeadAttentionLM", None), ("MultiHeadAttentionLM", None),
                         ("BlocksLM", None), ("ResidualBlocksLM", None), ("TransformerLM", 0.0),
                         ("TransformerLM", 0.1)]:
        params = dict(P)
        if p_drop is not None:
            params["dropout"] = p_drop
        m = build(kind, params)
        sd = torch.load(os.path.join(HERE, "checkpoints", kind + ".pt"), weights_only=True)
        m.load_state_dict(sd)
        m.train()
        torch.manual_seed(4242)
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, betas=(0.9, 0.95))
        losses = []
        for x, y in batches:
            _, loss = m(x, y)
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss))
        key = kind if not p_drop else f"{kind}_p{p_drop}"
        fin = m.state_dict()
        out[key] = {"losses": torch.tensor(losses, dtype=torch.float64),
                    "final_lm_or_tok": (fin.get("lm_head.weight", fin["token_embedding_table.weight"])).clone(),
                    "final_ln_f": fin["ln_f.weight"].clone() if "ln_f.weight" in fin else None}
    torch.save(out, os.path.join(HERE, "train_curves.pt"))


def modules():
    """Per-module forward + input/param grads at small shapes (reference init, seed 7)."""
    out = {}
    torch.manual_seed(7)
    C, T, NH, B = 32, 8, 4, 5
    specs = {
        "Head": lambda: ref_mc.Head(16, C, T),
        "MultiHeadAttention": lambda: ref_mc.MultiHeadAttention(NH, C // NH, C, T),
        "FeedForward": lambda: ref_mc.FeedForward(C),
        "Block": lambda: ref_mc.Block(C, T, NH),
        "FeedForward2": lambda: ref_mc.FeedForward2(C),
        "MultiHeadAttention2": lambda: ref_mc.MultiHeadAttention2(NH, C // NH, C, T),
        "ResidualBlock": lambda: ref_mc.ResidualBlock(C, NH, T),
        "FeedForward3": lambda: ref_mc.FeedForward3(C, 0.0),
        "Head2": lambda: ref_mc.Head2(16, C, T, 0.0),
        "MultiHeadAttention3": lambda: ref_mc.MultiHeadAttention3(NH, C // NH, C, T, 0.0),
        "ResidualBlock2": lambda: ref_mc.ResidualBlock2(C, NH, T, 0.0),
    }
    for name, mk in specs.items():
        m = mk()
        if name == "ResidualBlock2":  # make LN affine non-trivial
            with torch.no_grad():
                for ln in (m.ln1, m.ln2):
                    ln.weight.add_(0.1 * torch.randn(C))
                    ln.bias.add_(0.1 * torch.randn(C))
        recs = []
        for Tq in (T, 5):
            x = torch.randn(B, Tq, C, requires_grad=True)
            y = m(x)
            w = torch.randn_like(y)
            m.zero_grad()
            (y * w).sum().backward()
            recs.append({"x": x.detach().clone(), "y": y.detach().clone(), "w": w,
                         "dx": x.grad.clone(),
                         "grads": {k: p.grad.clone() for k, p in m.named_parameters()}})
        out[name] = {"state_dict": OrderedDict((k, v.clone()) for k, v in m.state_dict().items()), "cases": recs}
    torch.save(out, os.path.join(HERE, "module_vectors.pt"))


def scaled():
    """TransformerLM_scaled shape with formula weights (oracle.synthetic_state_dict)."""
    cfg = dict(vocab_size=80, embedding_dim=384, context_length=256, num_heads=6, num_layers=6)
    sd = O.synthetic_state_dict("TransformerLM", seed=20240, **cfg)
    m = ref_model.TransformerLM(80, 384, 256, 6, 6, 0.2)
    m.load_state_dict(sd, strict=True)
    m.eval()
    g = torch.Generator().manual_seed(99)
    x = torch.randint(0, 80, (4, 256), generator=g)
    y = torch.randint(0, 80, (4, 256), generator=g)
    m.zero_grad()
    lg, loss = m(x, y)
    loss.backward()
    rec = {"cfg": cfg, "seed": 20240, "idx": x, "targets": y, "loss": loss.detach().clone(),
           "logits_rows": lg.detach()[::37].clone(), "row_stride": 37,
           "grad_norms": {k: float(p.grad.norm()) for k, p in m.named_parameters() if p.grad is not None},
           "grad_lm_head": m.lm_head.weight.grad.clone(),
           "grad_qkv_l0h0_key": m.blocks[0].sa_head.heads[0].key.weight.grad.clone(),
           "grad_ffn_l5_b0": m.blocks[5].ffwd.net[0].bias.grad.clone(),
           "grad_ln1_l3_w": m.blocks[3].ln1.weight.grad.clone(),
           "grad_pos": m.position_embedding_table.weight.grad.clone(),
           "greedy_1": greedy(m, torch.zeros((1, 1), dtype=torch.long), 24),
           "n_params": sum(p.numel() for p in m.parameters())}
    torch.save(rec, os.path.join(HERE, "scaled_vectors.pt"))


SCALED_CURVE = dict(cfg=dict(vocab_size=80, embedding_dim=384, context_length=256, num_heads=6, num_layers=6),
                    init_seed=31415, batch_seed=27182, B=8, T=256, steps=200, lr=3e-4, betas=(0.9, 0.95))


def scaled_curves():
    """200 AdamW steps of the UNMODIFIED reference TransformerLM at the scaled shape (src/config.py:27-38,
    step glue src/train.py:143-151), B=8 x T=256 per step, from oracle.synthetic_state_dict init:
    one curve with dropout 0 and three torch seeds with dropout 0.2 (the seed-to-seed band)."""
    sc = SCALED_CURVE
    batches = O.structured_batches(sc["batch_seed"], sc["steps"], sc["B"], sc["T"])
    out = dict(sc)
    for p_drop, seeds in ((0.0, (0,)), (0.2, (1, 2, 3))):
        curves = []
        for sd_seed in seeds:
            sd = O.synthetic_state_dict("TransformerLM", seed=sc["init_seed"], **sc["cfg"])
            m = ref_model.TransformerLM(80, 384, 256, 6, 6, p_drop)
            m.load_state_dict(sd, strict=True)
            m.train()
            torch.manual_seed(sd_seed)
            opt = torch.optim.AdamW(m.parameters(), lr=sc["lr"], betas=sc["betas"])
            losses = []
            for x, y in batches:
                _, loss = m(x, y)
                opt.zero_grad()
                loss.backward()
                opt.step()
                losses.append(float(loss))
            curves.append(torch.tensor(losses, dtype=torch.float64))
            print("scaled curve p=%.1f seed %d: %.4f -> %.4f" % (p_drop, sd_seed, losses[0], losses[-1]), flush=True)
        out["p%.1f" % p_drop] = torch.stack(curves)
    torch.save(out, os.path.join(HERE, "scaled_curves.pt"))


def misc():
    out = {}
    # tokenizer on a synthetic corpus incl. non-ASCII
    text = "Started from the bottom now we're here\néü你好 ~ 0123456789!?,.;:'\"()[]-\n"
    enc, dec, vs = ref_pre.get_mapper(text)
    probe = "here we are 你 é!\n"
    out["tok"] = {"text": text, "vocab_size": vs, "probe": probe, "ids": enc(probe), "all_ids": enc(text)}
    assert dec(enc(text)) == text
    # get_batch windows
    torch.manual_seed(11)
    data = torch.arange(1000, dtype=torch.long) % 80
    xb, yb = ref_pre.get_batch(data, 8, 4, torch.device("cpu"))
    out["get_batch"] = {"seed": 11, "x": xb, "y": yb}
    # model_params estimator
    S = dict(context_length=256, embedding_dim=384, num_layers=6)
    out["model_params"] = {(k, s): ref_model.model_params(S if s else P, k, 80) for k in O.KINDS for s in (0, 1)}
    # CyclicLR trace
    w = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([w], lr=1e-3, betas=(0.9, 0.95))
    sch = torch.optim.lr_scheduler.CyclicLR(opt, base_lr=1e-3, max_lr=5e-3, step_size_up=5,
                                            mode="triangular", cycle_momentum=False)
    lrs = [opt.param_groups[0]["lr"]]
    for _ in range(20):
        opt.step()
        sch.step()
        lrs.append(opt.param_groups[0]["lr"])
    out["cyclic_lr"] = lrs
    # AdamW three-step known answer incl. a param without grad
    torch.manual_seed(3)
    a = torch.nn.Parameter(torch.randn(7, 5))
    b = torch.nn.Parameter(torch.randn(5))
    a0, b0 = a.detach().clone(), b.detach().clone()
    opt = torch.optim.AdamW([a, b], lr=3e-3, betas=(0.9, 0.95))
    gs = [torch.randn(7, 5) for _ in range(3)]
    for gi in gs:
        a.grad = gi.clone()
        b.grad = None
        opt.step()
    out["adamw"] = {"a0": a0, "b0": b0, "grads": gs, "a3": a.detach().clone(), "b3": b.detach().clone(), "lr": 3e-3}
    # actual param counts
    out["param_counts"] = {k: sum(p.numel() for p in build(k).parameters()) for k in O.KINDS}
    torch.save(out, os.path.join(HERE, "misc_vectors.pt"))


if __name__ == "__main__":
    if len(sys.argv) > 1:  # regenerate selected fixtures only: python make_golden.py scaled_curves ...
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    checkpoints()
    train_curves()
    modules()
    scaled()
    scaled_curves()
    misc()
    for f in sorted(os.listdir(HERE)):
        p = os.path.join(HERE, f)
        if os.path.isfile(p):
            print(f, os.path.getsize(p))
