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

    The sampler calls the network 2*num_steps-1 times per generated frame on identical shapes (edm2/sampler.py:53-75);
    each call is a few hundred tiny launches, i.e. launch-bound.  With the decode state made static
    (decode_state.make_static: fixed conv-context buffers, paged KV pools, device-side frame counters) nothing a launch
    depends on changes from frame to frame, so TWO graphs are captured once per (network, cache) -- "evaluate" and
    "evaluate and commit the frame to the cache" -- and replayed for every step of every generated frame.  They live
    on the cache dict, so successive edm_sampler_with_mse calls on the same cache reuse them."""

    def __init__(self, net, cache, conditioning):
        from .decode_state import cache_generation, make_static
        make_static(cache)
        self.net, self.cache = net, cache
        self.graphs = {}
        self.generation = cache_generation(cache)
        self.warm = False
        b = cache['shape'][0]
        self.cond = None if conditioning is None else conditioning.clone()

    def _eval(self, x, tt, update):
        return self.net(x, tt, self.cond, cache=self.cache, update_cache=update, just_2d=False)[0]

    def __call__(self, x, tt, conditioning, update_cache):
        from .decode_state import advance_host_counters, cache_generation, ensure_capacity
        ensure_capacity(self.cache, 1)
        if cache_generation(self.cache) != self.generation:      # a pool was re-allocated: old graphs point at freed pages
            self.graphs, self.generation = {}, cache_generation(self.cache)
        if self.cond is not None:
            self.cond.copy_(conditioning)
        if not self.warm:                        # eager warm-up (library handles, operand caches, rotary tables); a
            self.warm = True                     # non-committing evaluation has no side effects
            if not update_cache:
                return self._eval(x, tt, False)
            self._eval(x, tt, False)
        if update_cache not in self.graphs:
            sx, st = x.clone(), tt.clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # host mirrors advance while the committing evaluation is being RECORDED; undo that, the replay below is
            # the one execution
            with torch.cuda.graph(g):
                out = self._eval(sx, st, update_cache)
            if update_cache:
                advance_host_counters(self.cache, -1)
            self.graphs[update_cache] = (g, sx, st, out)
        g, sx, st, out = self.graphs[update_cache]
        sx.copy_(x)
        st.copy_(tt)
        g.replay()
        if update_cache:
            advance_host_counters(self.cache, 1)
        return out.clone()


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
    graphed = None
    if use_cuda_graph and guidance == 1 and device.type == "cuda":
        graphed = cache.get('_graphed_eval')
        if graphed is None or graphed.net is not net:
            graphed = cache['_graphed_eval'] = _GraphedEval(net, cache, conditioning)

    def denoise(x, t, cache, update_cache):
        tt = torch.ones(b, 1, device=device, dtype=dtype) * t
        if graphed is not None:
            return graphed(x, tt, conditioning, update_cache), cache
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
