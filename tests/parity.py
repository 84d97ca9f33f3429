"""Shared parity metric (BASELINE.json north_star: max rel err <= 2e-2, mean <= 2e-3 vs the fp32 reference).

Element-wise relative error is meaningless for tensors that cross zero, so:
  max_rel  = max|got - ref| / max|ref|
  mean_rel = mean|got - ref| / rms(ref)
"""
import os

import torch

MAX_REL = 2e-2
MEAN_REL = 2e-3


def errors(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    d = (got - ref).abs()
    return (d.max() / ref.abs().max().clamp_min(1e-30)).item(), (d.mean() / ref.pow(2).mean().sqrt().clamp_min(1e-30)).item()


def assert_close(got, ref, name="", max_rel=MAX_REL, mean_rel=MEAN_REL):
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    mx, mn = errors(got, ref)
    report = os.environ.get("ONIRIS_PARITY_REPORT")
    if report:     # evidence file: every comparison a test run made, with its measured errors and budget
        test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
        with open(report, "a") as f:
            f.write(f"{test}\t{name}\tmax_rel={mx:.3e}\tmean_rel={mn:.3e}\tbudget={max_rel:g}/{mean_rel:g}\n")
    assert mx <= max_rel and mn <= mean_rel, f"{name}: max_rel={mx:.3e} (<= {max_rel}) mean_rel={mn:.3e} (<= {mean_rel})"
    return mx, mn


def bf16r(x):
    """Round to the nearest bf16-representable fp32 value (so both sides see identical inputs)."""
    return x.to(torch.bfloat16).to(torch.float32)
