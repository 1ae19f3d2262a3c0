"""The two entry points end to end on the GPU (reference: src/train.py:85-183, src/inference.py:10-58), the
evaluation loop against the CPU oracle on the same windows, the device-side batch sampler, and the
training-state save / resume that upstream lacks (src/train.py:181-183 saves weights only)."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from drakegpt_b200 import inference, preprocessing, train  # noqa: E402
from drakegpt_b200 import model as M  # noqa: E402
from oracle import drake_oracle as O  # noqa: E402

DEV = "cuda"
SMALL = ["BigramLM", "SingleHeadAttentionLM", "MultiHeadAttentionLM", "BlocksLM", "ResidualBlocksLM", "TransformerLM"]


def test_train_and_inference_cli_scaled_transformer(tmp_path, capsys):
    """`train.py --model TransformerLM --scale true` (fused bf16 engine, CUDA-graph step, graph-captured eval),
    then `inference.py` on the checkpoint it wrote: 300 characters, i.e. past the 256-token window slide."""
    m = train.main(["--model", "TransformerLM", "--scale", "true", "--synthetic", "20000", "--iters", "20",
                    "--eval-interval", "10", "--eval-iters", "2", "--model-dir", str(tmp_path), "--generate", "24"])
    out = capsys.readouterr().out
    assert "Selected TransformerLM model for training. Model has" in out
    assert "step 10: train loss" in out and "step 20: train loss" in out
    assert "--- Predicting 24 characters with TransformerLM ---" in out
    assert m.precision == "bf16" and "_eval_graphs" in m.runner().__dict__  # eval ran through the captured graph
    path = tmp_path / "TransformerLM_scaled.pt"
    assert path.exists() and (tmp_path / "TransformerLM_scaled.pt.vocab.txt").exists()
    sd = torch.load(path, map_location="cpu", weights_only=True)
    assert len(sd) == 210 and sd["blocks.5.sa_head.heads.5.key.weight"].shape == (64, 384)
    assert all(v.dtype == torch.float32 and torch.isfinite(v).all() for v in sd.values())
    assert torch.equal(sd["ln_f.weight"], torch.ones(384))  # never applied, never trained, never decayed (Q1)
    # the training loss moved on the structured synthetic corpus
    l10 = float(out.split("step 10: train loss ")[1].split(",")[0])
    l20 = float(out.split("step 20: train loss ")[1].split(",")[0])
    assert l20 < l10 < 6.0
    ids, txt = inference.main(["--model", "TransformerLM", "--scale", "true", "--length", "300", "--model-dir", str(tmp_path),
                               "--out-dir", str(tmp_path / "inf")])
    assert len(ids) == 301 and os.path.exists(txt)
    vocab = (tmp_path / "TransformerLM_scaled.pt.vocab.txt").read_text(encoding="utf-8")
    text = open(txt, encoding="utf-8").read()
    assert len(text) == 301 and set(text) <= set(vocab)
    assert "Generating 300 character text using TransformerLM..." in capsys.readouterr().out


@pytest.mark.parametrize("kind", SMALL)
def test_train_and_inference_cli_small_models(kind, tmp_path, capsys):
    """Every model of the reference trains through the CLI at the PARAMS shape (exact fp32 kernels, autograd path,
    fused flat AdamW) and its checkpoint round-trips through inference.py."""
    train.main(["--model", kind, "--synthetic", "6000", "--iters", "8", "--eval-interval", "4", "--eval-iters", "2",
                "--model-dir", str(tmp_path), "--generate", "12"])
    out = capsys.readouterr().out
    assert "step 4: train loss" in out and "step 8: train loss" in out and f"saved {tmp_path}" in out
    ids, txt = inference.main(["--model", kind, "--length", "20", "--model-dir", str(tmp_path), "--out-dir", str(tmp_path)])
    assert len(ids) == 21 and len(glob.glob(str(tmp_path / f"generation_{kind}_*.txt"))) == 1


def test_inference_cli_on_shipped_checkpoint(tmp_path):
    """model/TransformerLM.pt as shipped by the reference (mps-tagged), greedy ids == the reference's."""
    from conftest import GOLDEN, load_golden
    ids, _ = inference.main(["--model", "TransformerLM", "--length", "64", "--greedy", "--model-dir",
                             os.path.join(GOLDEN, "checkpoints"), "--out-dir", str(tmp_path)])
    assert ids == load_golden("ckpt_vectors.pt")["TransformerLM"]["greedy_1"][0].tolist()


@pytest.mark.parametrize("precision,B,T,cfg,tol", [
    ("bf16", 8, 256, dict(vocab_size=80, embedding_dim=384, context_length=256, num_heads=6, num_layers=6), 3e-3),
    ("fp32", 32, 8, dict(vocab_size=80, embedding_dim=32, context_length=8, num_heads=4, num_layers=3), 2e-5)])
def test_evaluate_loss_equals_oracle_on_the_same_windows(precision, B, T, cfg, tol):
    """train.evaluate_loss (src/train.py:61-75) == the CPU oracle evaluated batch by batch on the windows the
    reference's get_batch draws from the same torch seed."""
    sd = O.synthetic_state_dict("TransformerLM", seed=77, **cfg)
    m = M.TransformerLM(cfg["vocab_size"], cfg["embedding_dim"], cfg["context_length"], cfg["num_heads"],
                        cfg["num_layers"], 0.2, precision=precision)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    g = torch.Generator().manual_seed(3)
    data = torch.randint(0, 80, (5000,), generator=g)
    tr, va = data[:4000], data[4000:]
    iters = 3
    torch.manual_seed(1234)
    got = train.evaluate_loss(tr, va, m, iters, T, B, torch.device(DEV))
    torch.manual_seed(1234)
    for name, split in (("train", tr), ("val", va)):
        ref = 0.0
        for _ in range(iters):
            x, y = preprocessing.get_batch(split, T, B, torch.device("cpu"))
            ref += float(O.forward("TransformerLM", sd, x, y)[1])
        ref /= iters
        assert got[name].device.type == "cpu" and got[name].dim() == 0
        assert abs(float(got[name]) - ref) <= tol * ref, (name, float(got[name]), ref)
    # a second call reuses the captured graph and sees parameter updates made in between
    if precision == "bf16":
        with torch.no_grad():
            m.lm_head.weight.mul_(0.5)  # a GEMM weight: the captured graph must see the re-cast bf16 shadow
        torch.manual_seed(1234)
        again = train.evaluate_loss(tr, va, m, iters, T, B, torch.device(DEV))
        assert abs(float(again["train"]) - float(got["train"])) > 1e-3


def test_device_batcher_windows():
    g = torch.Generator().manual_seed(1)
    data = torch.randint(0, 80, (3000,), generator=g)
    b1 = preprocessing.DeviceBatcher(data, 256, 64, torch.device(DEV), seed=7)
    b2 = preprocessing.DeviceBatcher(data, 256, 64, torch.device(DEV), seed=7)
    b3 = preprocessing.DeviceBatcher(data, 256, 64, torch.device(DEV), seed=8)
    seen = set()
    for _ in range(5):
        x, y = b1.next()
        x2, y2 = b2.next()
        x3, _ = b3.next()
        assert x.shape == (64, 256) and y.shape == (64, 256) and x.dtype == torch.int64 and x.is_cuda
        assert torch.equal(x, x2) and torch.equal(y, y2) and not torch.equal(x, x3)
        assert torch.equal(x[:, 1:], y[:, :-1])  # y is x shifted by one (src/preprocessing.py:43-45)
        xc, yc = x.cpu(), y.cpu()
        # every row is a real window of the corpus, starting inside [0, len - T)
        for r in range(0, 64, 9):
            starts = (data[: len(data) - 256] == xc[r, 0]).nonzero().view(-1).tolist()
            ok = [s for s in starts if torch.equal(data[s:s + 256], xc[r]) and torch.equal(data[s + 1:s + 257], yc[r])]
            assert ok, r
            seen.add(ok[0])
    assert len(seen) > 20  # windows differ from row to row and step to step


def test_training_state_roundtrip_and_resume(tmp_path, capsys):
    """Full training state (weights, Adam moments + step, CyclicLR position, iteration, dropout counter, sampler and
    torch RNG): save -> load restores every piece bit for bit, and `train 20` == `train 10, resume, train 10 more`
    up to the run-to-run summation-order noise of the atomically accumulated gradients (split-K wgrad, embedding
    and LayerNorm-parameter gradients use fp32 atomics, so even two straight runs differ in the last bits)."""
    common = ["--model", "TransformerLM", "--scale", "true", "--synthetic", "20000", "--eval-interval", "10",
              "--eval-iters", "1", "--generate", "0"]
    dA, dB = tmp_path / "a", tmp_path / "b"
    train.main(common + ["--iters", "10", "--checkpoint-every", "10", "--model-dir", str(dA), "--save", "false"])
    state_path = dA / "TransformerLM_scaled.state.pt"
    st = torch.load(state_path, map_location="cpu", weights_only=False)
    assert st["iter"] == 10 and st["adam_step"] == 10 and st["sched_steps"] == 1 and st["dropout_counter"] == 10
    assert abs(st["lr"] - train.cyclic_lr(1, 3e-4, 6e-4)) < 1e-12
    # bit-exact restore of every component into a fresh model
    m = M.TransformerLM(len(set(train.synthetic_corpus(20000))), 384, 256, 6, 6, 0.2).to(DEV)
    r = m.runner()
    opt = r.configure_optimizer(lr=1.0)
    bt = preprocessing.DeviceBatcher(torch.zeros(1000, dtype=torch.long), 256, 64, torch.device(DEV), seed=0)
    it, ss = train.load_training_state(st, m, r, opt, bt)
    back = train.training_state(m, r, opt, it, ss, bt)
    assert (it, ss) == (10, 1) and back["lr"] == st["lr"] and back["adam_step"] == 10
    for k in ("adam_m", "adam_v", "torch_rng", "batcher_rng"):
        assert torch.equal(back[k], st[k]), k
    assert all(torch.equal(back["model"][k], st["model"][k]) for k in st["model"])
    assert torch.equal(r.flat.shadow.float().cpu(), r.flat.p.bfloat16().float().cpu())  # bf16 shadow re-cast
    # resumed run vs straight run
    mA = train.main(common + ["--iters", "20", "--resume", str(state_path), "--model-dir", str(dA), "--save", "false"])
    outA = capsys.readouterr().out
    mB = train.main(common + ["--iters", "20", "--model-dir", str(dB), "--save", "false"])
    outB = capsys.readouterr().out
    assert "resumed" in outA and "step 20: train loss" in outA
    lA = float(outA.split("step 20: train loss ")[1].split(",")[0])
    lB = float(outB.split("step 20: train loss ")[1].split(",")[0])
    assert abs(lA - lB) <= 2e-3 * lB, (lA, lB)
    sdA, sdB, sd0 = mA.state_dict(), mB.state_dict(), st["model"]
    for k in ("lm_head.weight", "blocks.0.ffwd.net.0.weight", "blocks.5.sa_head.proj.weight", "token_embedding_table.weight"):
        upd = (sdB[k].cpu() - sd0[k]).norm()
        assert (sdA[k].cpu() - sdB[k].cpu()).norm() <= 2e-2 * upd, k  # same trajectory from iteration 10 on
    with pytest.raises(ValueError):
        train.load_training_state({"format": "something else"}, m, r, opt)
