"""torchrun --nproc-per-node N tools/check_dp_consistency.py: after two accumulation cycles with DIFFERENT data per rank,
every rank must hold bit-identical parameters (the data-parallel contract of cs_train.py's DDP) -- in the eager path, in
the CUDA-graph path (where the decoder's gradients are all-reduced behind an event recorded INSIDE the replayed graph,
concurrently with the rest of the backward pass), and with that overlap switched off; the three must agree with each other
to fp32 reduction-order noise."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200.train import LL_UNET, Trainer, init_distributed  # noqa: E402

rank, world, local = init_distributed()
dev = f"cuda:{local}"


def run(mode):
    os.environ["ONIRIS_NO_EARLY_REDUCE"] = "1" if mode == "eager, no overlap" else "0"
    tr = Trainer(LL_UNET, accumulation_steps=2, device=dev, seed=7)
    tr.unet.out_gain.data.fill_(1.0)   # random-init weights: keep the output path live
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    xs = [torch.randn(1, 4, 8, 64, 64, device=dev, generator=g) for _ in range(8)]
    if mode == "graphed":
        tr.capture(xs[0])              # two warm-up cycles on xs[0] inside
        for x in xs[4:]:
            tr.graphed_micro_step(x)
    else:
        for x in [xs[0]] * 4 + xs[4:]:
            tr.micro_step(x)
    torch.cuda.synchronize()
    ps = [p.detach().double() for p in tr.params if p.grad is not None]
    digest = torch.stack([sum(p.sum() for p in ps), sum(p.abs().sum() for p in ps), sum(p.reshape(-1)[::97].pow(2).sum() for p in ps)])
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    early = len(tr.buckets.early_groups)
    if rank == 0:
        print(f"world={world} [{mode}]: early-gradient groups {early}; parameter digests identical across ranks: {same}; "
              f"finite: {bool(torch.isfinite(digest).all())}; optimizer steps {tr.opt.opt_state[0].item():.0f}; digest {digest.tolist()}", flush=True)
    assert same and torch.isfinite(digest).all()
    return digest


d = {m: run(m) for m in ("eager", "eager, no overlap", "graphed")}
ref = d["eager, no overlap"]
for m, v in d.items():
    rel = float(((v - ref).abs() / ref.abs().clamp_min(1e-30)).max())
    if rank == 0:
        print(f"  {m}: max relative digest difference to 'eager, no overlap' {rel:.2e}", flush=True)
    assert rel < 1e-4
dist.destroy_process_group()
