"""Forward / backward split of one training micro-step (events on a parked stream), with and without the
weight-gradient side stream."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200.ops import WeightGradBranch  # noqa: E402
from autoregressive_diffusion_b200.train import CS_UNET, Trainer  # noqa: E402

WeightGradBranch.priority = int(os.environ.get("SIDE_PRIO", "0"))
WeightGradBranch.backward_pdl = int(os.environ.get("BWD_PDL", "0"))
main_stream = torch.cuda.Stream(priority=int(os.environ.get("MAIN_PRIO", "0")))
torch.cuda.set_stream(main_stream)
tr = Trainer(CS_UNET, device="cuda")
tr.unet.out_gain.data.fill_(1.0)   # random-init weights: keep the output path live
x = torch.randn(2, 16, 8, 32, 32, device="cuda")
for _ in range(6):
    tr.micro_step(x)


def measure():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    torch.cuda._sleep(int(1.5e9))
    ev[0].record()
    tr.micro += 1
    loss, _ = tr.loss_fn(tr.precond, x, None)
    ev[1].record()
    WeightGradBranch.defer_join = True
    (loss / tr.accum).backward()
    ev.append(torch.cuda.Event(enable_timing=True))
    ev[3].record()                       # main stream's own backward work done (branch not yet joined)
    WeightGradBranch.defer_join = False
    WeightGradBranch.join(x.device)
    ev[2].record()
    torch.cuda.synchronize()
    print(f"   main-stream backward chain {ev[1].elapsed_time(ev[3]):.2f} ms, then waits {ev[3].elapsed_time(ev[2]):.2f} ms for the weight-gradient stream")
    return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])


for enabled in (True, False, True, False):
    WeightGradBranch.enabled = enabled
    measure()
    f, b = measure()
    print(f"weight-gradient stream {'on ' if enabled else 'off'}: forward {f:.2f} ms  backward {b:.2f} ms  total {f + b:.2f} ms")
