"""Training entry point -- the reference's ``src/train.py`` on the B200 kernels.

    python -m drakegpt_b200.train --model TransformerLM --scale true --save true

Same flags, defaults, prints and checkpoint naming as the reference (src/train.py:85-183).
Documented differences:
  * ``--scale/--save`` parse true/false properly (the reference's ``type=bool`` treats any
    non-empty string as True);
  * with ``--scale`` the batch shape and learning rates come from SCALE_PARAMS (the reference
    always uses PARAMS, so its "scaled" run trains a 256-context model on 8-token windows);
    pass ``--reference-batching`` to reproduce the reference behaviour;
  * wandb is optional (``--wandb``; no network on the GPU boxes);
  * the optimizer is the fused flat AdamW and, for TransformerLM, the whole step is a CUDA graph;
    under ``torchrun`` the step is data-parallel (one process per GPU, one NCCL gradient all-reduce per
    step inside the captured graph; ``DGPT_DP_OVERLAP=1`` selects the bucketed, overlapped schedule);
  * ``--synthetic N`` trains on a synthetic N-character corpus when ../data is absent (the Kaggle
    corpus cannot be downloaded here).
"""
import argparse
import os

import torch

from . import config as cfg
from .model import (BigramLM, BlocksLM, MultiHeadAttentionLM, ResidualBlocksLM, SingleHeadAttentionLM,
                    TransformerLM, model_params)
from .preprocessing import DeviceBatcher, get_batch, get_mapper

MODEL_CLASSES = {
    "BigramLM": BigramLM, "SingleHeadAttentionLM": SingleHeadAttentionLM, "MultiHeadAttentionLM": MultiHeadAttentionLM,
    "BlocksLM": BlocksLM, "ResidualBlocksLM": ResidualBlocksLM, "TransformerLM": TransformerLM,
}


def _bool(s):
    if isinstance(s, bool):
        return s
    if s.lower() in ("1", "true", "yes", "y", "t"):
        return True
    if s.lower() in ("0", "false", "no", "n", "f", ""):
        return False
    raise argparse.ArgumentTypeError(f"expected a boolean, got {s!r}")


def build_model(model_name, scale, params, scale_params, vocab_size, device):
    """name -> (model on device, constructor kwargs, selected params); src/train.py:16-59."""
    if scale:
        params = scale_params
    if model_name not in MODEL_CLASSES:
        raise KeyError(f"unknown model {model_name!r}; choose from {sorted(MODEL_CLASSES)}")
    C, T = params["embedding_dim"], params["context_length"]
    model_config = {
        "BigramLM": {"vocab_size": vocab_size},
        "SingleHeadAttentionLM": dict(vocab_size=vocab_size, embedding_dim=C, context_length=T,
                                      head_size=params["head_size"]),
        "MultiHeadAttentionLM": dict(vocab_size=vocab_size, embedding_dim=C, context_length=T,
                                     head_size=params["head_size"], num_heads=params["num_heads"]),
        "BlocksLM": dict(vocab_size=vocab_size, embedding_dim=C, context_length=T, num_heads=params["num_heads"],
                         num_layers=params["num_layers"]),
        "ResidualBlocksLM": dict(vocab_size=vocab_size, embedding_dim=C, context_length=T,
                                 num_heads=params["num_heads"], num_layers=params["num_layers"]),
        "TransformerLM": dict(vocab_size=vocab_size, embedding_dim=C, context_length=T, num_heads=params["num_heads"],
                              num_layers=params["num_layers"], dropout=params["dropout"]),
    }[model_name]
    model = MODEL_CLASSES[model_name](**model_config).to(device)
    return model, model_config, params


@torch.no_grad()
def evaluate_loss(train_data, val_data, model, eval_iters, context_length, batch_size, device):
    """Mean loss over eval_iters random batches of each split (src/train.py:61-75).

    Losses stay on the device and are averaged there: one host sync per split instead of one per batch.
    """
    out = {}
    for name, data in (("train", train_data), ("val", val_data)):
        acc = torch.zeros((), device=device)
        for _ in range(eval_iters):
            x, y = get_batch(data, context_length, batch_size, device)
            _, loss = model(x, y)
            acc += loss
        out[name] = (acc / eval_iters).cpu()
    return out


def get_model_path(dir, model_name, scale):
    """<dir>/<model>[_scaled].pt (src/train.py:77-83)."""
    return os.path.join(dir, f"{model_name}_scaled.pt" if scale else f"{model_name}.pt")


def cyclic_lr(step, base_lr, max_lr, step_size_up=5):
    """CyclicLR(mode='triangular', step_size_up=5) value after `step` scheduler steps (src/train.py:122-126)."""
    import math
    total = 2 * step_size_up
    cycle = math.floor(1 + step / total)
    x = 1.0 + step / total - cycle
    ratio = step_size_up / total
    scale = x / ratio if x <= ratio else (x - 1) / (ratio - 1)
    return base_lr + (max_lr - base_lr) * scale


def synthetic_corpus(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    alphabet = "\n abcdefghijklmnopqrstuvwxyz',.!?ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789()-:;\"&*/[]"
    ids = torch.randint(0, len(alphabet), (n,), generator=g)
    ids[1::3] = (ids[::3][: len(ids[1::3])] * 5 + 1) % len(alphabet)
    return "".join(alphabet[i] for i in ids.tolist())


def main(argv=None):
    from ._lib import require_gpu
    from .parallel import init_from_env
    parser = argparse.ArgumentParser(description="Train a language model")
    parser.add_argument("--model", type=str, default="TransformerLM", help="Model to train")
    parser.add_argument("--scale", type=_bool, default=False, help="Train scaled model")
    parser.add_argument("--save", type=_bool, default=True, help="Save model")
    parser.add_argument("--iters", type=int, default=cfg.TRAIN["iters"])
    parser.add_argument("--eval-interval", type=int, default=cfg.TRAIN["eval_interval"])
    parser.add_argument("--eval-iters", type=int, default=cfg.TRAIN["eval_iters"])
    parser.add_argument("--reference-batching", action="store_true", help="always batch with PARAMS (reference quirk)")
    parser.add_argument("--synthetic", type=int, default=0, help="train on a synthetic corpus of N characters")
    parser.add_argument("--wandb", action="store_true")
    args = parser.parse_args(argv)

    torch.manual_seed(42)
    rank, world, local = init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("drakegpt_b200 needs a B200 GPU: there is no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    require_gpu()
    if rank == 0:
        print(f"Using device: {device}")

    if args.synthetic or not os.path.exists(cfg.DATA["input"]):
        text = synthetic_corpus(args.synthetic or 200000)
        encode, decode, vocab_size = get_mapper(text)
        data = torch.tensor(encode(text), dtype=torch.long)
        n = int(0.9 * len(data))
        train_data, val_data = data[:n], data[n:]
    else:
        with open(cfg.DATA["input"], "r", encoding="utf-8") as f:
            text = f.read()
        train_data = torch.load(cfg.DATA["train"], map_location="cpu")
        val_data = torch.load(cfg.DATA["val"], map_location="cpu")
        encode, decode, vocab_size = get_mapper(text)

    model, model_config, params = build_model(args.model, args.scale, cfg.PARAMS, cfg.SCALE_PARAMS, vocab_size, device)
    if rank == 0:
        print(f"Selected {args.model} model for training. Model has {model_params(params, args.model, vocab_size)} parameters.")
    hp = cfg.PARAMS if args.reference_batching else params
    T, B = hp["context_length"], hp["batch_size"]
    model.train()

    fused = args.model == "TransformerLM"
    sched_steps = 0
    if fused:
        from .graph import GraphedTrainStep
        runner = model.runner()
        runner.base_seed = 42 + rank
        opt = runner.configure_optimizer(lr=hp["base_lr"], betas=hp["betas"])
        reducer = runner.make_reducer() if world > 1 else None
        step = GraphedTrainStep(runner, B, T, reducer)
        batcher = DeviceBatcher(train_data, T, B, device, seed=42 + rank)
    else:
        from .optim import FlatParams, FusedAdamW
        flat = FlatParams(model)
        flat.attach_grads()
        opt = FusedAdamW(flat, lr=hp["base_lr"], betas=hp["betas"])

    run = None
    if args.wandb and rank == 0:
        import wandb
        wandb.login()
        model_config.update(scheduler="CyclicLR", learning_rate=hp["base_lr"], betas=hp["betas"], batch_size=B)
        run = wandb.init(project="DrakeGPT", config=model_config, name=args.model)

    if rank == 0:
        print(f"--- Training {args.model} ---")
    for it in range(args.iters):
        if fused:
            x, y = batcher.next()
            step.step(x, y)
        else:
            x, y = get_batch(train_data, T, B, device)
            _, loss = model(x, y)
            loss.backward()
            opt.step()
        if (it + 1) % args.eval_interval == 0:
            model.eval()
            losses = evaluate_loss(train_data, val_data, model, args.eval_iters, T, B, device)
            sched_steps += 1  # the reference steps CyclicLR only here
            opt.param_groups[0]["lr"] = cyclic_lr(sched_steps, hp["base_lr"], hp["max_lr"])
            opt.upload()
            if run is not None:
                run.log({"train_loss": losses["train"], "val_loss": losses["val"]})
            if rank == 0:
                print(f"step {it + 1}: train loss {losses['train']:.4f}, val loss {losses['val']:.4f}")
            model.train()

    if rank == 0:
        print(f"--- Predicting 100 characters with {args.model} ---")
        model.eval()
        idx = torch.zeros((1, 1), dtype=torch.long, device=device)
        print(decode(model.generate(idx, max_new_tokens=100)[0].tolist()))
        if args.save:
            os.makedirs(cfg.MODEL_DIR, exist_ok=True)
            path = get_model_path(cfg.MODEL_DIR, args.model, args.scale)
            torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)
            with open(path + ".vocab.txt", "w", encoding="utf-8") as f:  # tokenizer sidecar
                f.write("".join(sorted(set(text))))
            print(f"saved {path}")


if __name__ == "__main__":
    main()
