"""The training driver end to end: CUDA-graph replay with the weight-gradient stream, PDL and the fused optimizer must
reproduce the plain eager single-stream run (same data, noise levels and noise) to fp32 reduction-order noise."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SMALL_UNET = dict(img_resolution=16, img_channels=8, label_dim=4, model_channels=64, channel_mult=[1, 2],
                  channel_mult_noise=None, channel_mult_emb=None, num_blocks=1, video_attn_resolutions=[8],
                  frame_attn_resolutions=[16])


class FixedNoiseLoss:
    """EDM2Loss with the noise levels and the noise supplied by the test, so two runs see identical inputs."""

    def __init__(self, base, sigma, noise):
        self.base, self.sigma, self.noise = base, sigma, noise

    def __call__(self, net, images, conditioning=None, just_2d=False):
        return self.base(net, images, conditioning, sigma=self.sigma, noise=self.noise, just_2d=just_2d)


def _run(graphed, x, sigma, noise, steps):
    from autoregressive_diffusion_b200 import _lib
    from autoregressive_diffusion_b200.ops import WeightGradBranch
    from autoregressive_diffusion_b200.train import Trainer
    WeightGradBranch.enabled = graphed
    prev = _lib.query("ob_set_pdl", 1 if graphed else 0)
    try:
        tr = Trainer(SMALL_UNET, accumulation_steps=2, lr=1e-2, device="cuda", seed=3)
        tr.loss_fn = FixedNoiseLoss(tr.loss_fn, sigma, noise)
        start = {id(p): p.detach().clone() for p in tr.params}
        losses = []
        if graphed:
            tr.capture(x)          # warm-up cycles inside capture() also step the optimizer: replay from that state
        else:
            for _ in range(4):     # the same 2 warm-up cycles capture() runs
                tr.micro_step(x)
        for _ in range(steps):
            out = tr.graphed_micro_step(x) if graphed else tr.micro_step(x)[0]
            losses.append(float(out))
        torch.cuda.synchronize()
        moved = max(float((p.detach() - start[id(p)]).abs().max()) for p in tr.params if p.grad is not None)
        return torch.cat([p.detach().reshape(-1) for p in tr.params]), losses, moved, float(tr.opt.step_lr[0])
    finally:
        WeightGradBranch.enabled = True
        _lib.query("ob_set_pdl", prev)


def test_graph_replay_matches_eager_training():
    torch.manual_seed(11)
    b, n = 2, 4
    x = torch.randn(b, n, 8, 16, 16, device="cuda")
    sigma = torch.cat((torch.rand(b, 1, device="cuda").expand(-1, n) * 0.1, (torch.randn(b, n, device="cuda") + 0.9).exp()), dim=1)
    noise = torch.randn(b, 2 * n, 8, 16, 16, device="cuda")
    p_eager, l_eager, moved_e, steps_e = _run(False, x, sigma, noise, steps=4)
    p_graph, l_graph, moved_g, steps_g = _run(True, x, sigma, noise, steps=4)
    assert steps_e == steps_g == 4.0                       # 2 warm-up cycles + 2 measured cycles of 2 micro-steps
    assert moved_e > 1e-3 and moved_g > 1e-3               # the optimizer really moved the weights
    assert all(torch.isfinite(torch.tensor(l_eager + l_graph)))
    for a, c in zip(l_eager, l_graph):
        assert abs(a - c) <= 2e-2 * abs(a), (l_eager, l_graph)
    diff = (p_eager - p_graph).abs()
    scale = p_eager.abs().mean()
    # Adam divides by sqrt(v): elements with a near-zero gradient amplify reduction-order noise, so bound the bulk tightly
    # and the tail loosely
    assert float(diff.mean()) <= 2e-3 * float(scale), (float(diff.mean()), float(scale))
    assert float(diff.quantile(0.999)) <= 5e-2 * float(scale) + 2e-2, float(diff.quantile(0.999))
