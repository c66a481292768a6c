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


# ---------------------------------------------------------------------------------------------------------------------
# Neural-spline flow (the reference's default flow='spline'): autograd side.  Sampling runs in the CUDA kernels of
# csrc/nnb_spline.cu with parameters exported by Trainer._sync_device(); this module is what Adam differentiates.
# Parameter names / shapes are those of nnest/networks.py:393-705 so that models/netG.pt stays interchangeable:
#   flow.flows.<3k>.{s,t}                 ActNorm (data-dependent initialisation on the first forward batch)
#   flow.flows.<3k+1>.{L,S,U}             Invertible1x1Conv, W = P L (U + diag S); P is fixed and NOT in the state_dict
#   flow.flows.<3k+2>.{f1,f2}.net.{0,2,4,6}.{weight,bias}     NSF_CL conditioners (Linear, LeakyReLU(0.2) x 3, Linear)
# The rational-quadratic spline is evaluated densely (every element, clamped to the tail bound, selected with where())
# instead of on a boolean selection: same values, static shapes -- the training step can be captured in a CUDA graph.
# ---------------------------------------------------------------------------------------------------------------------
import torch.nn.functional as F

_MIN_BIN, _MIN_DERIV = 1e-3, 1e-3


def _knots(unnorm, bound):
    """cumulative knot positions (..., K + 1) on [-bound, bound] from unnormalised widths / heights (RQS, networks.py:479-497)"""
    K = unnorm.shape[-1]
    w = _MIN_BIN + (1 - _MIN_BIN * K) * F.softmax(unnorm, dim=-1)
    cum = 2 * bound * torch.cumsum(w, dim=-1)[..., :-1] - bound
    edge = torch.full_like(unnorm[..., :1], bound)
    return torch.cat([-edge, cum, edge], dim=-1)


def _rqs_dense(inputs, W, H, D, inverse, bound):
    """unconstrained_RQS + RQS (networks.py:431-553) on every element; identity outside [-bound, bound]."""
    inside = (inputs >= -bound) & (inputs <= bound)
    v = inputs.clamp(-bound, bound)
    const = math.log(math.exp(1 - _MIN_DERIV) - 1)
    deriv = _MIN_DERIV + F.softplus(F.pad(D, (1, 1), value=const))
    cumw, cumh = _knots(W, bound), _knots(H, bound)
    widths, heights = cumw[..., 1:] - cumw[..., :-1], cumh[..., 1:] - cumh[..., :-1]
    K = widths.shape[-1]
    knots = (cumh if inverse else cumw).detach().clone()
    knots[..., -1] += 1e-6                                     # searchsorted's eps (networks.py:425-430)
    idx = ((v[..., None] >= knots).sum(dim=-1) - 1).clamp(0, K - 1)[..., None]
    g = lambda t: t.gather(-1, idx)[..., 0]
    in_cw, in_w, in_ch, in_h = g(cumw), g(widths), g(cumh), g(heights)
    delta = g(heights / widths)
    d0, d1 = g(deriv), g(deriv[..., 1:])
    if inverse:
        yv = v - in_ch
        a = yv * (d0 + d1 - 2 * delta) + in_h * (delta - d0)
        b = in_h * d0 - yv * (d0 + d1 - 2 * delta)
        c = -delta * yv
        disc = (b * b - 4 * a * c).clamp_min(0)
        root = (2 * c) / (-b - torch.sqrt(disc))
        out = root * in_w + in_cw
        tt = root * (1 - root)
        den = delta + (d0 + d1 - 2 * delta) * tt
        num = delta ** 2 * (d1 * root ** 2 + 2 * delta * tt + d0 * (1 - root) ** 2)
        ld = -(torch.log(num) - 2 * torch.log(den))
    else:
        theta = (v - in_cw) / in_w
        tt = theta * (1 - theta)
        num = in_h * (delta * theta ** 2 + d0 * tt)
        den = delta + (d0 + d1 - 2 * delta) * tt
        out = in_ch + num / den
        dnum = delta ** 2 * (d1 * theta ** 2 + 2 * delta * tt + d0 * (1 - theta) ** 2)
        ld = torch.log(dnum) - 2 * torch.log(den)
    return torch.where(inside, out, inputs), torch.where(inside, ld, torch.zeros_like(ld))


class _SplineMLP(nn.Module):

    def __init__(self, nin, nout, nh):
        super(_SplineMLP, self).__init__()
        self.net = nn.Sequential(nn.Linear(nin, nh), nn.LeakyReLU(0.2), nn.Linear(nh, nh), nn.LeakyReLU(0.2),
                                 nn.Linear(nh, nh), nn.LeakyReLU(0.2), nn.Linear(nh, nout))

    def forward(self, x):
        return self.net(x)


class SplineCoupling(nn.Module):
    """NSF_CL (networks.py:556-619): upper | lower, then lower | new upper; uneven split for odd dimensions."""

    def __init__(self, dim, K=8, B=3, hidden_dim=8):
        super(SplineCoupling, self).__init__()
        self.dim, self.K, self.B = dim, K, B
        self.nlow = dim // 2 + (dim & 1)
        self.nup = dim - self.nlow
        self.f1 = _SplineMLP(self.nlow, (3 * K - 1) * self.nup, hidden_dim)
        self.f2 = _SplineMLP(self.nup, (3 * K - 1) * self.nlow, hidden_dim)

    def _params(self, net, cond, m):
        out = net(cond).reshape(-1, m, 3 * self.K - 1)
        W, H, D = out[..., :self.K], out[..., self.K:2 * self.K], out[..., 2 * self.K:]
        return 2 * self.B * F.softmax(W, dim=2), 2 * self.B * F.softmax(H, dim=2), F.softplus(D)

    def forward(self, x):
        lower, upper = x[:, :self.nlow], x[:, self.nlow:]
        upper, l1 = _rqs_dense(upper, *self._params(self.f1, lower, self.nup), False, self.B)
        lower, l2 = _rqs_dense(lower, *self._params(self.f2, upper, self.nlow), False, self.B)
        return torch.cat([lower, upper], dim=1), l1.sum(dim=1) + l2.sum(dim=1)

    def inverse(self, z):
        lower, upper = z[:, :self.nlow], z[:, self.nlow:]
        lower, l1 = _rqs_dense(lower, *self._params(self.f2, upper, self.nlow), True, self.B)
        upper, l2 = _rqs_dense(upper, *self._params(self.f1, lower, self.nup), True, self.B)
        return torch.cat([lower, upper], dim=1), l1.sum(dim=1) + l2.sum(dim=1)


class Conv1x1(nn.Module):
    """Invertible1x1Conv (networks.py:622-653): LU-parametrised, P fixed (a non-persistent buffer: not in the state_dict)."""

    def __init__(self, dim):
        super(Conv1x1, self).__init__()
        self.dim = dim
        Q = torch.nn.init.orthogonal_(torch.randn(dim, dim))
        P, L, U = torch.linalg.lu(Q)
        self.register_buffer('P', P, persistent=False)
        self.L = nn.Parameter(L)
        self.S = nn.Parameter(U.diag().clone())
        self.U = nn.Parameter(torch.triu(U, diagonal=1))

    def assemble(self):
        eye = torch.eye(self.dim, device=self.L.device, dtype=self.L.dtype)
        L = torch.tril(self.L, diagonal=-1) + eye
        U = torch.triu(self.U, diagonal=1)
        return self.P @ L @ (U + torch.diag(self.S))

    def forward(self, x):
        return x @ self.assemble(), torch.sum(torch.log(torch.abs(self.S))).expand(x.shape[0])

    def inverse(self, z):
        return z @ torch.inverse(self.assemble()), (-torch.sum(torch.log(torch.abs(self.S)))).expand(z.shape[0])


class ActNorm(nn.Module):
    """ActNorm (networks.py:656-695): z = x exp(s) + t, initialised from the first batch seen by forward()."""

    def __init__(self, dim):
        super(ActNorm, self).__init__()
        self.s = nn.Parameter(torch.randn(1, dim))
        self.t = nn.Parameter(torch.randn(1, dim))
        self.data_dep_init_done = False

    def forward(self, x):
        if not self.data_dep_init_done:
            with torch.no_grad():
                self.s.copy_(-torch.log(x.std(dim=0, keepdim=True)))
                self.t.copy_(-(x * torch.exp(self.s)).mean(dim=0, keepdim=True))
            self.data_dep_init_done = True
        return x * torch.exp(self.s) + self.t, torch.sum(self.s, dim=1).expand(x.shape[0])

    def inverse(self, z):
        return (z - self.t) * torch.exp(-self.s), torch.sum(-self.s, dim=1).expand(z.shape[0])


class SingleSpeedSpline(nn.Module):
    """SingleSpeedSpline (networks.py:698-705) = [ActNorm, Invertible1x1Conv, NSF_CL] x num_blocks."""

    def __init__(self, num_inputs, hidden_dim, num_blocks, num_bins=8, tail_bound=3, prior=None, device=None):
        super(SingleSpeedSpline, self).__init__()
        self.num_inputs, self.num_hidden, self.num_blocks = num_inputs, hidden_dim, num_blocks
        self.num_bins, self.tail_bound = num_bins, tail_bound
        layers = []
        for _ in range(num_blocks):
            layers += [ActNorm(num_inputs), Conv1x1(num_inputs),
                       SplineCoupling(num_inputs, K=num_bins, B=tail_bound, hidden_dim=hidden_dim)]
        self.flow = _Stack(layers)
        self.device = device
        if device is not None:
            self.flow.to(device)
        self._std_normal = prior is None
        if prior is None:
            loc = torch.zeros(num_inputs, device=device)
            prior = torch.distributions.MultivariateNormal(loc, torch.eye(num_inputs, device=device))
        self.prior = prior

    forward = SingleSpeedNVP.forward
    inverse = SingleSpeedNVP.inverse
    log_probs = SingleSpeedNVP.log_probs
    sample = SingleSpeedNVP.sample

    def needs_data_init(self):
        return any(isinstance(m, ActNorm) and not m.data_dep_init_done for m in self.flow.flows)

    def mark_initialised(self):
        for m in self.flow.flows:
            if isinstance(m, ActNorm):
                m.data_dep_init_done = True

    def set_permutations(self, mats):
        """fixed permutation matrices of the 1x1 convolutions (not part of the state_dict), one per block"""
        convs = [m for m in self.flow.flows if isinstance(m, Conv1x1)]
        assert len(mats) == len(convs)
        with torch.no_grad():
            for m, P in zip(convs, mats):
                m.P.copy_(torch.as_tensor(P, dtype=m.P.dtype, device=m.P.device))

    def packed_for_kernel(self):
        """flat float32 vector in the layout nnb_set_flow_spline documents (include/nnb.h)"""
        parts = []
        with torch.no_grad():
            fl = list(self.flow.flows)
            for k in range(self.num_blocks):
                an, cv, cl = fl[3 * k], fl[3 * k + 1], fl[3 * k + 2]
                W = cv.assemble().float()
                Winv = torch.linalg.inv(W.double()).float()
                parts += [an.s.reshape(-1), an.t.reshape(-1), W.reshape(-1), Winv.reshape(-1),
                          torch.sum(torch.log(torch.abs(cv.S))).reshape(1)]
                for net in (cl.f1, cl.f2):
                    for j in (0, 2, 4, 6):
                        parts += [net.net[j].weight.reshape(-1), net.net[j].bias.reshape(-1)]
            return torch.cat([p.float().reshape(-1) for p in parts]).cpu().numpy()
