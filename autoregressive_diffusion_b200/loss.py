"""EDM2 training loss on the clean (+) noised DART sequence -- same call signature as edm2/loss.py:9-47."""
import torch


class EDM2Loss:
    def __init__(self, P_mean=0.5, P_std=2., sigma_data=1., context_noise_reduction=0.1):
        assert 0 <= context_noise_reduction <= 1, f"context_noise_reduction must be in [0,1], got {context_noise_reduction}"
        self.P_mean, self.P_std, self.sigma_data = P_mean, P_std, sigma_data
        self.context_noise_reduction = context_noise_reduction

    def draw_sigma(self, batch_size, n_frames, device, just_2d=False):
        """Noise levels (edm2/loss.py:24-28): log-normal for the targets, one small uniform level per sequence for the context."""
        sigma = (torch.randn(batch_size, n_frames, device=device) * self.P_std + self.P_mean).exp()
        if just_2d:
            return sigma
        ctx = torch.rand(batch_size, 1, device=device).expand(-1, n_frames) * self.context_noise_reduction
        return torch.cat((ctx, sigma), dim=1)

    def __call__(self, net, images, conditioning=None, sigma=None, just_2d=False, noise=None):
        b, n = images.shape[:2]
        assert net.training, "The model should be in training mode"
        seq = images if just_2d else torch.cat((images, images), dim=1)
        if conditioning is not None and not just_2d:
            conditioning = torch.cat((conditioning, conditioning), dim=1)
        if sigma is None:
            sigma = self.draw_sigma(b, n, images.device, just_2d)
        if noise is None:
            noise = torch.randn_like(seq)
        out, _ = net(seq + sigma.reshape(*sigma.shape, 1, 1, 1) * noise, sigma, conditioning, just_2d=just_2d)
        err = ((out[:, -n:] - images) ** 2).mean(dim=(-1, -2, -3))
        s = sigma[:, -n:]
        losses = err * (s ** 2 + self.sigma_data ** 2) / (s * self.sigma_data) ** 2
        unweighted = losses.mean().detach()
        net.noise_weight.add_data(s, losses)
        losses = losses / net.noise_weight.calculate_mean_loss(s)
        return losses.mean(), unweighted
