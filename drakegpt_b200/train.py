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
    corpus cannot be downloaded here);
  * ``--checkpoint-every N`` / ``--resume PATH`` save and restore the full training state (weights, Adam
    moments and step, CyclicLR position, iteration, dropout counter, sampler RNG) -- absent upstream, which
    saves only ``state_dict`` at the end (src/train.py:181-183).
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

    Same batches as the reference (``get_batch`` consumes the torch CPU generator identically).  Losses stay
    on the device: one host sync per split instead of one ``loss.item()`` per batch, and for the fused
    TransformerLM engine the eval forward + loss accumulation is one CUDA-graph replay per batch
    (``graph.GraphedEvalStep``).
    """
    out = {}
    graphed = None
    if isinstance(model, TransformerLM) and model.precision == "bf16" and not model.training:
        from .graph import GraphedEvalStep
        runner = model.runner()
        runner._reattach()
        cache = runner.__dict__.setdefault("_eval_graphs", {})
        key = (batch_size, context_length, runner._flat_gen)
        graphed = cache.get(key)
        if graphed is None:
            graphed = cache[key] = GraphedEvalStep(runner, batch_size, context_length)
    for name, data in (("train", train_data), ("val", val_data)):
        if graphed is not None:
            graphed.reset()
            for _ in range(eval_iters):
                x, y = get_batch(data, context_length, batch_size, device)
                graphed.step(x, y)
            out[name] = (graphed.loss_sum / eval_iters).cpu()
            continue
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


# --------------------------------------------------------------------------- #
# full training state (absent upstream: src/train.py:181-183 saves the weights only)
# --------------------------------------------------------------------------- #
def training_state(model, runner, opt, it, sched_steps, batcher=None, reducer=None):
    """Everything a bit-faithful continuation needs, as CPU tensors / plain values.

    Under data parallelism with the fused peer-memory optimizer (``parallel.PeerAdamW``) every rank holds the Adam
    moments of its own shard only: pass the ``reducer`` and call this on EVERY rank (it gathers them)."""
    flat = runner.flat
    if reducer is not None and getattr(reducer, "fused_optimizer", False):
        import torch.distributed as dist
        dist.barrier()            # no rank is inside a step any more
        reducer.sync_master()     # fp32 masters owned by other ranks (only their bf16 shadows are broadcast per step)
        adam_m, adam_v = reducer.gather_moments()
    else:
        adam_m, adam_v = flat.m.detach().cpu().clone(), flat.v.detach().cpu().clone()
    st = {
        "format": "drakegpt_b200.training_state.v1",
        "model": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
        "adam_m": adam_m, "adam_v": adam_v,
        "adam_step": int(opt.step_dev.item()), "lr": float(opt.param_groups[0]["lr"]),
        "sched_steps": int(sched_steps), "iter": int(it),
        "dropout_counter": int(runner.seed_dev.item()), "base_seed": int(runner.base_seed),
        "torch_rng": torch.get_rng_state(),
        "slots": {n: (o, k) for n, (o, k, _) in flat.slots.items()},
    }
    if batcher is not None:
        st["batcher_rng"] = batcher.gen.get_state().cpu()
    return st


def load_training_state(st, model, runner, opt, batcher=None):
    """Inverse of ``training_state``; returns (next iteration, scheduler steps taken)."""
    if st.get("format") != "drakegpt_b200.training_state.v1":
        raise ValueError("not a drakegpt_b200 training-state file")
    model.load_state_dict(st["model"], strict=True)
    runner._reattach()
    flat = runner.flat
    if {n: (o, k) for n, (o, k, _) in flat.slots.items()} != st["slots"]:
        raise ValueError("training state was saved from a different model layout")
    flat.m.copy_(st["adam_m"])
    flat.v.copy_(st["adam_v"])
    flat.g.zero_()
    flat.refresh_shadow(force=True)
    opt.step_dev.fill_(st["adam_step"])
    opt.t = st["adam_step"]
    opt.param_groups[0]["lr"] = st["lr"]
    opt.upload()
    runner.seed_dev.fill_(st["dropout_counter"])
    runner.base_seed = st["base_seed"]
    torch.set_rng_state(st["torch_rng"])
    if batcher is not None and "batcher_rng" in st:
        batcher.gen.set_state(st["batcher_rng"])
    return st["iter"], st["sched_steps"]


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
    parser.add_argument("--model-dir", type=str, default=str(cfg.MODEL_DIR), help="where checkpoints are written")
    parser.add_argument("--checkpoint-every", type=int, default=0, help="save the full training state every N iters")
    parser.add_argument("--resume", type=str, default="", help="training-state file to continue from")
    parser.add_argument("--generate", type=int, default=100, help="characters to sample after training")
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

    # ONE flat parameter arena per model: the Runner's (generation, the fused step and the optimizer all use it)
    fused = args.model == "TransformerLM" and model.precision == "bf16"
    runner = model.runner()
    runner.base_seed = 42 + rank
    opt = runner.configure_optimizer(lr=hp["base_lr"], betas=hp["betas"])
    flat = runner.flat
    batcher, step, reducer = None, None, None
    if fused:
        from .graph import GraphedTrainStep
        reducer = runner.make_reducer() if world > 1 else None
        step = GraphedTrainStep(runner, B, T, reducer)
        batcher = DeviceBatcher(train_data, T, B, device, seed=42 + rank)
    elif world > 1:
        import torch.distributed as dist
        opt.grad_scale = 1.0 / world  # the SUM all-reduce below, averaged inside the fused AdamW

    start_it, sched_steps = 0, 0
    if args.resume:
        st = torch.load(args.resume, map_location="cpu", weights_only=False)
        start_it, sched_steps = load_training_state(st, model, runner, opt, batcher)
        if rank == 0:
            print(f"resumed {args.resume} at iteration {start_it}")

    run = None
    if args.wandb and rank == 0:
        import wandb
        wandb.login()
        model_config.update(scheduler="CyclicLR", learning_rate=hp["base_lr"], betas=hp["betas"], batch_size=B)
        run = wandb.init(project="DrakeGPT", config=model_config, name=args.model)

    name = f"{args.model}_scaled" if args.scale else args.model
    if rank == 0:
        print(f"--- Training {args.model} ---")
    for it in range(start_it, args.iters):
        if fused:
            x, y = batcher.next()
            step.step(x, y)
        else:
            x, y = get_batch(train_data, T, B, device)
            _, loss = model(x, y)
            loss.backward()
            if world > 1:  # data parallel on the autograd path: SUM the flat gradient arena across ranks
                dist.all_reduce(flat.g[:flat.n_live], op=dist.ReduceOp.SUM)
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
        if args.checkpoint_every and (it + 1) % args.checkpoint_every == 0:
            st = training_state(model, runner, opt, it + 1, sched_steps, batcher, reducer)  # collective under DP
            if rank == 0:
                os.makedirs(args.model_dir, exist_ok=True)
                path = os.path.join(args.model_dir, f"{name}.state.pt")
                torch.save(st, path)
                print(f"saved training state {path} at iteration {it + 1}")

    if reducer is not None and getattr(reducer, "fused_optimizer", False):
        import torch.distributed as dist
        dist.barrier()
        reducer.sync_master()  # before state_dict(): fetch the fp32 masters this rank does not own
    if rank == 0:
        model.eval()
        if args.generate > 0:
            print(f"--- Predicting {args.generate} characters with {args.model} ---")
            idx = torch.zeros((1, 1), dtype=torch.long, device=device)
            print(decode(model.generate(idx, max_new_tokens=args.generate)[0].tolist()))
        if args.save:
            os.makedirs(args.model_dir, exist_ok=True)
            path = get_model_path(args.model_dir, args.model, args.scale)
            torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)
            with open(path + ".vocab.txt", "w", encoding="utf-8") as f:  # tokenizer sidecar
                f.write("".join(sorted(set(text))))
            print(f"saved {path}")
    return model


if __name__ == "__main__":
    main()
