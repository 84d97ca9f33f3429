"""Per-shape time of every tensor-core launch in one training micro-step (events on a parked stream)."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200 import _lib  # noqa: E402
from autoregressive_diffusion_b200.ops import WeightGradBranch  # noqa: E402
from autoregressive_diffusion_b200.train import CS_UNET, Trainer  # noqa: E402


class Prof:
    names = {"ob_conv_fwd": 8, "ob_conv_fwd_fused": 8, "ob_conv_dgrad": 7, "ob_conv_wgrad": 5, "ob_conv_wgrad_acc": 5}

    def __init__(self):
        self.ev = []

    def before(self, name, args):
        if name in self.names:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        return None

    def after(self, name, args, e0):
        if e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            off = self.names[name]
            self.ev.append((name, tuple(args[off:off + 9]), e0, e1))


tr = Trainer(CS_UNET, device="cuda")
tr.unet.out_gain.data.fill_(1.0)   # random-init weights: keep the output path live
x = torch.randn(2, 16, 8, 32, 32, device="cuda")
for _ in range(6):
    tr.micro_step(x)
p = Prof()
WeightGradBranch.enabled = False     # time every launch alone
_lib.set_profiler(p)
torch.cuda.synchronize()
torch.cuda._sleep(int(6e8))
tr.micro_step(x)
torch.cuda.synchronize()
_lib.set_profiler(None)
agg = collections.defaultdict(lambda: [0, 0.0])
for name, shp, e0, e1 in p.ev:
    agg[(name, shp)][0] += 1
    agg[(name, shp)][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"total tensor-core launch time {tot:.2f} ms")
print("kernel           n_seq S  T   H   W  cin cout k g   count   ms_total  us_each  TFLOP/s")
for (name, shp), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    n_seq, S, T, H, W, cin, cout, k, gated = shp
    px = n_seq * T * H * W
    fl = 2.0 * px * S * cin * cout * 9 + 2.0 * px * cin * cout * 18 if gated else 2.0 * px * S * cin * cout * k * k
    print(f"{name:14s} {n_seq:4d} {S} {T:3d} {H:3d} {W:3d} {cin:4d} {cout:4d} {k} {gated}  {n:4d}  {ms:8.3f}  {1e3 * ms / n:7.1f}  {fl * n / ms / 1e9:7.1f}")
