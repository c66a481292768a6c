"""CPU oracle: neural-spline flow of the reference (its DEFAULT flow, flow='spline'), restated in numpy float32.

TEST INFRASTRUCTURE ONLY -- groundwork for SURVEY section 8(f) #3 (the CUDA kernels for this flow are not built yet; the
product package raises NotImplementedError for flow='spline').  Only tests/ may import this module.

Restates (no code shared with) the reference:
  * nnest/networks.py:401-417  MLP: Linear, LeakyReLU(0.2) x 3, Linear
  * nnest/networks.py:425-553  searchsorted, unconstrained_RQS, RQS (rational-quadratic spline, Durkan et al. 2019), including
                               the reference's quirks: W / H arrive already soft-maxed and scaled by 2B from NSF_CL and are
                               soft-maxed AGAIN inside RQS; the last knot is nudged by eps = 1e-6 for the bin search only;
                               boundary derivatives are padded so that min_derivative + softplus(.) = 1; identity outside the
                               tail bound
  * nnest/networks.py:556-619  NSF_CL (coupling: upper | lower, then lower | upper; uneven split for odd dimensions)
  * nnest/networks.py:622-653  Invertible1x1Conv (W = P L (U + diag S); log-det = sum log|S|; P is NOT in the state_dict)
  * nnest/networks.py:656-695  AffineConstantFlow / ActNorm (after its data-dependent initialisation)
  * nnest/networks.py:698-705  SingleSpeedSpline: [ActNorm, Invertible1x1Conv, NSF_CL] x num_blocks, num_bins = 8, tail bound 3
Parity is pinned against the real reference: tests/golden/make_golden_spline.py, tests/test_oracle_spline.py.
"""
import numpy as np

F32 = np.float32
MIN_BIN_WIDTH = MIN_BIN_HEIGHT = MIN_DERIVATIVE = 1e-3


def _softmax(a):
    e = np.exp(a - a.max(axis=-1, keepdims=True))
    return (e / e.sum(axis=-1, keepdims=True)).astype(F32)


def _softplus(a):
    # torch.nn.functional.softplus: log1p(exp(x)), linear above the threshold 20
    a = np.asarray(a, dtype=F32)
    return np.where(a > 20, a, np.log1p(np.exp(np.minimum(a, 20)))).astype(F32)


def _mlp(layers, x):
    h = x
    for i, (W, b) in enumerate(layers):
        h = (h @ W.T + b).astype(F32)
        if i + 1 < len(layers):
            h = np.where(h > 0, h, F32(0.2) * h).astype(F32)
    return h


def _rqs(inputs, uw, uh, ud, inverse, bound):
    """RQS on points strictly inside [-bound, bound]; uw / uh (n, K), ud (n, K + 1) unnormalised."""
    K = uw.shape[-1]
    left = bottom = F32(-bound)
    right = top = F32(bound)
    widths = _softmax(uw)
    widths = (F32(MIN_BIN_WIDTH) + F32(1 - MIN_BIN_WIDTH * K) * widths).astype(F32)
    cumw = np.concatenate([np.zeros((len(inputs), 1), F32), np.cumsum(widths, axis=-1, dtype=F32)], axis=-1)
    cumw = ((right - left) * cumw + left).astype(F32)
    cumw[:, 0], cumw[:, -1] = left, right
    widths = cumw[:, 1:] - cumw[:, :-1]
    deriv = (F32(MIN_DERIVATIVE) + _softplus(ud)).astype(F32)
    heights = _softmax(uh)
    heights = (F32(MIN_BIN_HEIGHT) + F32(1 - MIN_BIN_HEIGHT * K) * heights).astype(F32)
    cumh = np.concatenate([np.zeros((len(inputs), 1), F32), np.cumsum(heights, axis=-1, dtype=F32)], axis=-1)
    cumh = ((top - bottom) * cumh + bottom).astype(F32)
    cumh[:, 0], cumh[:, -1] = bottom, top
    heights = cumh[:, 1:] - cumh[:, :-1]
    knots = (cumh if inverse else cumw).copy()
    knots[:, -1] += F32(1e-6)                                     # searchsorted's eps, used for the search only
    idx = (inputs[:, None] >= knots).sum(axis=-1) - 1
    rows = np.arange(len(inputs))
    in_cw, in_w = cumw[rows, idx], widths[rows, idx]
    in_ch, in_h = cumh[rows, idx], heights[rows, idx]
    delta = (heights / widths).astype(F32)
    in_delta = delta[rows, idx]
    d0, d1 = deriv[rows, idx], deriv[rows, idx + 1]
    if inverse:
        y = (inputs - in_ch).astype(F32)
        a = y * (d0 + d1 - 2 * in_delta) + in_h * (in_delta - d0)
        b = in_h * d0 - y * (d0 + d1 - 2 * in_delta)
        c = -in_delta * y
        disc = b * b - 4 * a * c
        assert (disc >= 0).all()
        root = ((2 * c) / (-b - np.sqrt(disc))).astype(F32)
        out = (root * in_w + in_cw).astype(F32)
        tt = root * (1 - root)
        den = in_delta + (d0 + d1 - 2 * in_delta) * tt
        num = in_delta ** 2 * (d1 * root ** 2 + 2 * in_delta * tt + d0 * (1 - root) ** 2)
        return out, (-(np.log(num) - 2 * np.log(den))).astype(F32)
    theta = ((inputs - in_cw) / in_w).astype(F32)
    tt = theta * (1 - theta)
    num = in_h * (in_delta * theta ** 2 + d0 * tt)
    den = in_delta + (d0 + d1 - 2 * in_delta) * tt
    out = (in_ch + num / den).astype(F32)
    dnum = in_delta ** 2 * (d1 * theta ** 2 + 2 * in_delta * tt + d0 * (1 - theta) ** 2)
    return out, (np.log(dnum) - 2 * np.log(den)).astype(F32)


def _unconstrained_rqs(inputs, W, H, D, inverse, bound):
    """inputs (n, m); W, H (n, m, K); D (n, m, K - 1).  Identity (log-det 0) outside [-bound, bound]."""
    inside = (inputs >= -bound) & (inputs <= bound)
    out = inputs.astype(F32).copy()
    ld = np.zeros_like(out)
    const = F32(np.log(np.exp(1 - MIN_DERIVATIVE) - 1))
    Dp = np.concatenate([np.full(D.shape[:-1] + (1,), const, F32), D, np.full(D.shape[:-1] + (1,), const, F32)], axis=-1)
    if not inside.any():
        # the reference calls RQS on the (empty) selection, which raises (networks.py:464-465); Sampler._mcmc_sample
        # catches the ValueError and skips the proposal (sampler.py:320-324).  Only reachable when EVERY coordinate of
        # EVERY sample of the batch lies outside the tail bound, i.e. in practice for one-sample batches
        raise ValueError('No input values')
    out[inside], ld[inside] = _rqs(inputs[inside], W[inside], H[inside], Dp[inside], inverse, bound)
    return out, ld


class SplineWeights(object):
    """Parameters of a SingleSpeedSpline: per block ActNorm (s, t), 1x1 conv (P, L, S, U), NSF_CL (f1, f2 as lists of (W, b))."""

    def __init__(self, d, blocks, num_bins=8, tail_bound=3):
        self.d, self.blocks, self.K, self.B = int(d), blocks, int(num_bins), float(tail_bound)
        self.half = self.d // 2
        self.even = self.d == 2 * self.half

    @classmethod
    def from_golden(cls, g):
        d, nb = int(g['d']), int(g['blocks'])
        sd = {k[3:]: np.asarray(g[k], dtype=F32) for k in g.files if k.startswith('sd/')}
        blocks = []
        for k in range(nb):
            a, c, n = 3 * k, 3 * k + 1, 3 * k + 2
            mlp = lambda f: [(sd['flow.flows.%d.%s.net.%d.weight' % (n, f, j)], sd['flow.flows.%d.%s.net.%d.bias' % (n, f, j)])
                             for j in (0, 2, 4, 6)]
            blocks.append(dict(s=sd['flow.flows.%d.s' % a], t=sd['flow.flows.%d.t' % a], P=np.asarray(g['P/%d' % k], F32),
                               L=sd['flow.flows.%d.L' % c], S=sd['flow.flows.%d.S' % c], U=sd['flow.flows.%d.U' % c],
                               f1=mlp('f1'), f2=mlp('f2')))
        return cls(d, blocks, int(g['num_bins']), float(g['tail_bound']))

    def conv_matrix(self, blk):
        d = self.d
        L = np.tril(blk['L'], -1) + np.eye(d, dtype=F32)
        U = np.triu(blk['U'], 1) + np.diag(blk['S'])
        return (blk['P'] @ L @ U).astype(F32)


def _split_params(out, m, K, B):
    out = out.reshape(-1, m, 3 * K - 1)
    W, H, D = out[..., :K], out[..., K:2 * K], out[..., 2 * K:]
    return (2 * B * _softmax(W)).astype(F32), (2 * B * _softmax(H)).astype(F32), _softplus(D)


def _coupling(w, blk, x, inverse):
    """NSF_CL.forward / inverse (networks.py:573-619)."""
    h, K, B = w.half, w.K, w.B
    nlow = h if w.even else h + 1
    lower, upper = x[:, :nlow].astype(F32), x[:, nlow:].astype(F32)
    ld = np.zeros(x.shape[0], F32)
    if not inverse:
        W, H, D = _split_params(_mlp(blk['f1'], lower), h, K, B)
        upper, l1 = _unconstrained_rqs(upper, W, H, D, False, B)
        W, H, D = _split_params(_mlp(blk['f2'], upper), nlow, K, B)
        lower, l2 = _unconstrained_rqs(lower, W, H, D, False, B)
    else:
        W, H, D = _split_params(_mlp(blk['f2'], upper), nlow, K, B)
        lower, l1 = _unconstrained_rqs(lower, W, H, D, True, B)
        W, H, D = _split_params(_mlp(blk['f1'], lower), h, K, B)
        upper, l2 = _unconstrained_rqs(upper, W, H, D, True, B)
    ld = ld + l1.sum(axis=1) + l2.sum(axis=1)
    return np.concatenate([lower, upper], axis=1), ld.astype(F32)


def flow_forward(w, x):
    """netG.forward (networks.py:24-32 over SingleSpeedSpline's layer list): x -> (z, log|det dz/dx|)."""
    x = np.asarray(x, dtype=F32)
    ld = np.zeros(x.shape[0], F32)
    for blk in w.blocks:
        x = (x * np.exp(blk['s']) + blk['t']).astype(F32)                       # ActNorm
        ld = ld + blk['s'].sum()
        x = (x @ w.conv_matrix(blk)).astype(F32)                               # Invertible1x1Conv
        ld = ld + np.log(np.abs(blk['S'])).sum()
        x, l = _coupling(w, blk, x, False)
        ld = ld + l
    return x, ld.astype(F32)


def flow_inverse(w, z):
    """netG.inverse (networks.py:34-42): z -> (x, log|det dx/dz|)."""
    z = np.asarray(z, dtype=F32)
    ld = np.zeros(z.shape[0], F32)
    for blk in reversed(w.blocks):
        z, l = _coupling(w, blk, z, True)
        ld = ld + l
        z = (z @ np.linalg.inv(w.conv_matrix(blk)).astype(F32)).astype(F32)
        ld = ld - np.log(np.abs(blk['S'])).sum()
        z = ((z - blk['t']) * np.exp(-blk['s'])).astype(F32)
        ld = ld - blk['s'].sum()
    return z, ld.astype(F32)


def pack_for_kernel(w, hidden):
    """Flat float32 parameter vector in the layout of nnest_b200/csrc/nnb_spline.cuh (ActNorm s, t; assembled 1x1
    convolution matrix, its inverse (computed in float64) and log-det; the two conditioner MLPs), block after block."""
    parts = []
    for blk in w.blocks:
        Wc = w.conv_matrix(blk)
        Wci = np.linalg.inv(Wc.astype(np.float64)).astype(F32)
        parts += [blk['s'].ravel(), blk['t'].ravel(), Wc.ravel(), Wci.ravel(),
                  np.array([np.log(np.abs(blk['S'])).sum()], dtype=F32)]
        for f in ('f1', 'f2'):
            for W, b in blk[f]:
                parts += [W.ravel(), b.ravel()]
    return np.concatenate([np.asarray(a, dtype=F32) for a in parts])
