"""torchrun helper for tests/test_gpu_dp.py: 2-rank data-parallel training == 1-rank training on the
concatenated batch (dropout off), through the bucketed NCCL all-reduce inside the captured graph."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drakegpt_b200 import model as M  # noqa: E402
from drakegpt_b200.graph import GraphedTrainStep  # noqa: E402
from drakegpt_b200.parallel import init_from_env  # noqa: E402
from oracle import drake_oracle as O  # noqa: E402


def run(model_sd, batches, world, rank, graphed):
    dev = torch.device("cuda", torch.cuda.current_device())
    m = M.TransformerLM(80, 128, 128, 2, 2, 0.0, precision="bf16")
    m.load_state_dict(model_sd)
    m = m.to(dev).train()
    r = m.runner()
    r.configure_optimizer(lr=1e-3)
    red = r.make_reducer() if world > 1 else None
    B, T = batches[0][0].shape
    losses = []
    step = GraphedTrainStep(r, B, T, red) if graphed else None
    for x, y in batches:
        if step is not None:
            losses.append(step.step(x.to(dev), y.to(dev)).clone())
        else:
            losses.append(r.train_step(x.to(dev), y.to(dev), red).clone())
    torch.cuda.synchronize()
    if red is not None and getattr(red, "fused_optimizer", False):
        assert red.status() == 0, f"peer barrier timed out (status {red.status()})"
        lazy = red.need32 is not None
        if lazy:  # fp32 masters of the GEMM weights are only current on their owner until sync_master()
            assert red.master_stale
            dist.barrier()
            red.sync_master()
            assert not red.master_stale
            assert torch.equal(r.flat.shadow[:r.flat.n_live].float(), r.flat.p[:r.flat.n_live].bfloat16().float())
        mom_m, mom_v = red.gather_moments()  # full-length moments assembled from the shards
        assert mom_m.shape == (r.flat.n_live,) and torch.isfinite(mom_v).all() and float(mom_v.abs().sum()) > 0
        red.close()
    return torch.stack(losses).cpu(), {k: v.detach().cpu() for k, v in m.state_dict().items() if not k.endswith("tril")}


def main():
    rank, world, local = init_from_env("nccl")
    torch.cuda.set_device(local)
    cfg = dict(vocab_size=80, embedding_dim=128, context_length=128, num_heads=2, num_layers=2)
    sd = O.synthetic_state_dict("TransformerLM", seed=3, **cfg)
    g = torch.Generator().manual_seed(9)
    full = [(torch.randint(0, 80, (8, 128), generator=g), torch.randint(0, 80, (8, 128), generator=g)) for _ in range(4)]
    per = 8 // world
    mine = [(x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per]) for x, y in full]
    for graphed in (False, True):
        losses, params = run(sd, mine, world, rank, graphed)
        # every rank ends with identical parameters
        for k, v in params.items():
            ref = v.cuda()
            dist.broadcast(ref, 0)
            assert torch.equal(ref.cpu(), v), (k, "replicas diverged")
        mean_loss = losses.cuda()
        dist.all_reduce(mean_loss)
        mean_loss = (mean_loss / world).cpu()
        if rank == 0:
            l1, p1 = run(sd, full, 1, 0, graphed)  # the same global batch on one GPU
            assert ((mean_loss - l1).abs() / l1).max() < 5e-3, (mean_loss, l1)
            for k in p1:
                num = (params[k] - p1[k]).norm()
                den = (p1[k] - sd[k]).norm() + 1e-12  # compare the UPDATE, not the weights
                assert num / den < 0.1, (k, float(num / den))
    dist.barrier()
    if rank == 0:
        print(f"DP_CHECK_OK mode={os.environ.get('DGPT_DP_MODE', 'peer')}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
