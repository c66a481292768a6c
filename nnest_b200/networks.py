"""Trainable parameter container for the RealNVP flow (autograd side only).

The hot path never runs this module: sampling goes through the CUDA kernels with weights exported by
Trainer._sync_device().  It exists so that flow fitting (Adam on -log p, reference nnest/trainer.py:384-403)
keeps PyTorch autograd for now, and so that `models/netG.pt` stays interchangeable with the reference:
parameter names are those of nnest/networks.py (flow.flows.<k>.{scale_net,translate_net}.<2j>.{weight,bias},
ScaleLayer 'scale'), coupling masks are (arange(d) + k) % 2 and are not parameters.
"""
import math

import torch
import torch.nn as nn


def _mlp(n_in, n_hidden, n_layers, act):
    mods = [nn.Linear(n_in, n_hidden), act()]
    for _ in range(n_layers):
        mods += [nn.Linear(n_hidden, n_hidden), act()]
    mods.append(nn.Linear(n_hidden, n_in))
    return nn.Sequential(*mods)


class AffineCoupling(nn.Module):
    """y = x * exp(s(x_m)) + t(x_m) on the un-masked half (tanh s-net, relu t-net)."""

    def __init__(self, dim, hidden, n_layers, parity, translate_only=False):
        super(AffineCoupling, self).__init__()
        self.translate_only = translate_only
        self.register_buffer('keep', ((torch.arange(dim) + parity) % 2).float(), persistent=False)
        if not translate_only:
            self.scale_net = _mlp(dim, hidden, n_layers, nn.Tanh)
        self.translate_net = _mlp(dim, hidden, n_layers, nn.ReLU)

    def _st(self, v):
        vm = v * self.keep
        free = 1.0 - self.keep
        t = self.translate_net(vm) * free
        s = None if self.translate_only else self.scale_net(vm) * free
        return s, t

    def forward(self, x):
        s, t = self._st(x)
        if s is None:
            return x + t, x.new_zeros(x.shape[0])
        return x * torch.exp(s) + t, s.sum(-1)

    def inverse(self, z):
        s, t = self._st(z)
        if s is None:
            return z - t, z.new_zeros(z.shape[0])
        return (z - t) * torch.exp(-s), -s.sum(-1)


class ConstScale(nn.Module):
    """Learned global rescaling used when scale='constant' (log-det contribution is the scalar itself,
    as in the reference's ScaleLayer)."""

    def __init__(self):
        super(ConstScale, self).__init__()
        self.scale = nn.Parameter(torch.tensor(0.0))

    def forward(self, x):
        return x * torch.exp(self.scale), self.scale.expand(x.shape[0])

    def inverse(self, z):
        return z * torch.exp(-self.scale), (-self.scale).expand(z.shape[0])


class _Stack(nn.Module):

    def __init__(self, layers):
        super(_Stack, self).__init__()
        self.flows = nn.ModuleList(layers)

    def forward(self, x):
        total = x.new_zeros(x.shape[0])
        for layer in self.flows:
            x, ld = layer.forward(x)
            total = total + ld
        return x, total

    def inverse(self, z):
        total = z.new_zeros(z.shape[0])
        for layer in reversed(self.flows):
            z, ld = layer.inverse(z)
            total = total + ld
        return z, total


class SingleSpeedNVP(nn.Module):

    def __init__(self, num_inputs, num_hidden, num_blocks, num_layers, scale='', prior=None, device=None):
        super(SingleSpeedNVP, self).__init__()
        self.num_inputs, self.num_hidden = num_inputs, num_hidden
        self.num_blocks, self.num_layers, self.scale = num_blocks, num_layers, scale
        translate_only = scale in ('translate', 'constant')
        layers = []
        for k in range(num_blocks):
            layers.append(AffineCoupling(num_inputs, num_hidden, num_layers, k, translate_only))
            if scale == 'constant':
                layers.append(ConstScale())
        self.flow = _Stack(layers)
        self.device = device
        if device is not None:
            self.flow.to(device)
        self._std_normal = prior is None
        if prior is None:
            loc = torch.zeros(num_inputs, device=device)
            prior = torch.distributions.MultivariateNormal(loc, torch.eye(num_inputs, device=device))
        self.prior = prior

    def forward(self, x):
        return self.flow.forward(x)

    def inverse(self, z):
        return self.flow.inverse(z)

    def log_probs(self, inputs):
        u, log_det = self.forward(inputs)
        if self._std_normal:
            # N(0, I) log-density written out (what MultivariateNormal.log_prob evaluates); no host sync, so the
            # training step can be captured in a CUDA graph
            return -0.5 * (self.num_inputs * math.log(2 * math.pi) + (u * u).sum(-1)) + log_det
        lp = self.prior.log_prob(u)
        if lp.dim() > 1:
            lp = lp.sum(1)
        return lp + log_det

    def sample(self, num_samples=None, noise=None):
        if noise is None:
            noise = self.prior.sample((num_samples,))
        if self.device is not None:
            noise = noise.to(self.device)
        return self.inverse(noise)[0]
