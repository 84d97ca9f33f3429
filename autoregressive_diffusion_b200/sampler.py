"""Autoregressive next-frame sampler: the Heun/EDM loop of edm2/sampler.py:13-85 driving the cached decode path."""
import numpy as np
import torch


def sigma_schedule(num_steps, sigma_min, sigma_max, rho, device, dtype=torch.float32):
    """Karras schedule with the trailing zero (edm2/sampler.py:35-38)."""
    i = torch.arange(num_steps, dtype=dtype, device=device)
    t = (sigma_max ** (1 / rho) + i / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([t, torch.zeros_like(t[:1])])


class _GraphedEval:
    """CUDA-graph replay of the cached one-frame denoiser evaluation.

    Within one generated frame the sampler calls the network 2*num_steps-1 times on identical shapes against an
    unchanged cache; each call is ~500 tiny launches, i.e. launch-bound.  The second call is captured, the rest are
    replays with the noisy frame and sigma copied into static buffers.  The cache tensors are baked in by address, so a
    new graph is captured for every generated frame (the cache objects are replaced when it is updated)."""

    def __init__(self, net, conditioning):
        self.net, self.conditioning = net, conditioning
        self.graph, self.calls = None, 0

    def __call__(self, x, tt, cache):
        self.calls += 1
        if self.calls == 1:                      # eager warm-up (library handles, operand caches, rotary tables)
            return self.net(x, tt, self.conditioning, cache=cache, update_cache=False, just_2d=False)[0]
        if self.graph is None:
            self.sx, self.st = x.clone(), tt.clone()
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self.net(self.sx, self.st, self.conditioning, cache=cache, update_cache=False, just_2d=False)[0]
        self.sx.copy_(x)
        self.st.copy_(tt)
        self.graph.replay()
        return self.out.clone()


@torch.no_grad()
def edm_sampler_with_mse(net, cache, target=None, gnet=None, conditioning=None, num_steps=32, sigma_min=0.002, sigma_max=80,
                         rho=7, guidance=1, S_churn=0, S_min=0, S_max=float('inf'), S_noise=1, dtype=torch.float32,
                         x_init=None, use_cuda_graph=False):
    """Generate ONE frame after the frames summarised in `cache`.  Returns (frame, mse, mse_pred, cache) like the reference;
    `x_init` (unit-variance noise [b,1,c,h,w]) lets a caller fix the randomness; `use_cuda_graph` replays the
    non-updating network evaluations from a CUDA graph (guidance == 1 only)."""
    was_training = net.training
    net.eval()
    b, _, c, h, w = cache.get('shape', (None,) * 5)
    device = net.device
    graphed = _GraphedEval(net, conditioning) if (use_cuda_graph and guidance == 1 and device.type == "cuda") else None

    def denoise(x, t, cache, update_cache):
        tt = torch.ones(b, 1, device=device, dtype=dtype) * t
        if graphed is not None and not update_cache:
            shape = cache['shape']
            out = graphed(x, tt, cache)
            cache['shape'] = shape
            return out, cache
        dx, cache = net(x, tt, conditioning, cache=cache, update_cache=update_cache, just_2d=False)
        if guidance == 1:
            return dx, cache
        ref, _ = net(x, tt, conditioning, just_2d=True)
        return ref.lerp(dx, guidance), cache

    ts = sigma_schedule(num_steps, sigma_min, sigma_max, rho, device, dtype)
    if x_init is None:
        x_init = torch.randn(b, 1, c, h, w, device=device)
    x_next = x_init * ts[0]
    mse, mse_pred = [], []
    if target is not None:
        target = target.to(dtype)
        x_next = x_next + target
    for i, (t_cur, t_next) in enumerate(zip(ts[:-1], ts[1:])):
        x_hat, t_hat = x_next, t_cur
        if S_churn > 0 and S_min <= t_cur <= S_max:
            gamma = min(S_churn / num_steps, np.sqrt(2) - 1)
            t_hat = t_cur + gamma * t_cur
            x_hat = x_next + (t_hat ** 2 - t_cur ** 2).sqrt() * S_noise * torch.randn_like(x_next)
        # only the last Euler step commits its activations / keys to the cache (edm2/sampler.py:66)
        x_pred, cache = denoise(x_hat, t_hat, cache, update_cache=(i == num_steps - 1 and target is None))
        d_cur = (x_hat - x_pred) / t_hat
        x_next = x_hat + (t_next - t_hat) * d_cur
        if i < num_steps - 1:
            x_pred, _ = denoise(x_next, t_next, cache, update_cache=False)
            d_prime = (x_next - x_pred) / t_next
            x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
        if target is not None:
            mse_pred.append(torch.mean((x_pred - target) ** 2).item())
            mse.append(torch.mean((x_next - target) ** 2).item())
    if was_training:
        net.train()
    return x_next, mse, mse_pred, cache
