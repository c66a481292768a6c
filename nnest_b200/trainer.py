"""Flow facade + fitting with the reference's Trainer interface (nnest/trainer.py:28-301).

The facade methods the MCMC loop calls -- forward / inverse / get_samples / get_latent_samples /
get_prior_samples / get_synthetic_samples / log_probs (trainer.py:247-301) -- run on the hand-written
CUDA flow kernels through the C ABI (nnb_flow_forward / nnb_flow_inverse); there is no ATen path for
them.  Fitting (train / _train / _validate, trainer.py:134-245,384-418) keeps the reference's procedure
(nearest-neighbour jitter, 90/10 split, Adam, early stopping with patience, best-model restore, netG.pt /
originals.npy / tensorboard artefacts).  One epoch -- every mini-batch's forward pass, backward pass and Adam
step, plus the validation loss -- is ONE launch of the fused kernel behind nnb_train_epoch (csrc/nnb_train.cuh);
architectures that kernel does not cover (scale != '', hidden_dim 64, more than two hidden layers) are fitted with
PyTorch autograd captured in a CUDA graph.  After every change of the weights they are re-exported to the sampling
kernels (`_sync_device`) and broadcast to all ranks when multi-GPU.

Only flow='nvp' with num_slow=0 is implemented on the device (the hot path north_star
names); other flows raise NotImplementedError.
"""
import copy
import logging
import math
import os
import time

import numpy as np
import torch

from . import dist
from .engine import Engine
from .networks import SingleSpeedNVP, SingleSpeedSpline
from .utils.logger import create_logger


def _copy_into_params(params, vec):
    """In-place parameter update from a flat vector.  (torch's vector_to_parameters re-points .data at views of the
    vector, which would silently detach the parameters from the addresses a captured CUDA graph updates.)"""
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(vec[off:off + n].view_as(p))
            off += n


class _GraphedStep(object):
    """One Adam step on -mean(log p(x + jitter * eps)) captured in a CUDA graph: the flow is tiny (<= 12k parameters),
    so the eager iteration is pure launch latency (~100 kernels); replaying the graph is a single launch."""

    def __init__(self, net, optimizer, batch_size, dim, device):
        self.x = torch.zeros((batch_size, dim), device=device)
        self.jitter = torch.zeros((), device=device)
        self.loss = torch.zeros((), device=device)
        self.net, self.opt = net, optimizer

        def step():
            data = self.x + self.jitter * torch.randn_like(self.x)
            self.opt.zero_grad(set_to_none=False)
            loss = -self.net.log_probs(data).mean()
            loss.backward()
            self.opt.step()
            self.loss.copy_(loss.detach())

        # Warm-up (on a side stream) and capture must not change the weights / optimizer state training starts from.
        # Everything is restored IN PLACE: the graph keeps the addresses of the parameter and Adam state tensors.
        saved_w = torch.nn.utils.parameters_to_vector(list(net.parameters())).detach().clone()
        saved_state = {p: {k: v.clone() for k, v in st.items() if torch.is_tensor(v)}
                       for p, st in optimizer.state.items()}

        def restore():
            with torch.no_grad():
                _copy_into_params(list(net.parameters()), saved_w)
                for p, st in optimizer.state.items():
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            if p in saved_state and k in saved_state[p]:
                                v.copy_(saved_state[p][k])
                            else:
                                v.zero_()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        restore()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            step()
        restore()

    def __call__(self, batch, jitter):
        self.x.copy_(batch)
        self.jitter.fill_(float(jitter))
        self.graph.replay()
        return self.loss


class Trainer(object):
    best_validation_epoch = None
    best_validation_loss = None

    def __init__(self,
                 x_dim,
                 hidden_dim=16,
                 num_slow=0,
                 batch_size=100,
                 flow='spline',
                 scale='',
                 num_blocks=3,
                 num_layers=1,
                 base_dist=None,
                 load_model='',
                 log_dir='logs/test',
                 use_gpu=True,
                 log=True,
                 learning_rate=0.0001,
                 weight_decay=1e-6,
                 log_level=logging.INFO,
                 engine=None):
        if flow.lower() not in ('nvp', 'spline'):
            raise NotImplementedError("nnest_b200 implements flow='nvp' (RealNVP) and flow='spline' (neural spline "
                                      "flow, the reference's default); got flow=%r" % flow)
        self.flow_kind = flow.lower()
        if num_slow != 0:
            raise NotImplementedError('fast/slow flows (num_slow > 0) are not on the accelerated path')
        if base_dist is not None:
            raise NotImplementedError('only the default N(0, I) base distribution is supported')
        if not torch.cuda.is_available():
            raise RuntimeError('nnest_b200 needs a CUDA device (B200); there is no CPU fallback')

        self.engine = engine if engine is not None else Engine()
        self.device = self.engine.device
        self.x_dim = x_dim
        self.z_dim = x_dim
        self.batch_size = batch_size
        self.total_iters = 0
        self.num_slow = 0
        self.scale = scale

        if self.flow_kind == 'spline':      # trainer.py:92-98: 8 bins, tail bound 3; num_layers / scale do not apply
            self.netG = SingleSpeedSpline(x_dim, hidden_dim, num_blocks, num_bins=8, tail_bound=3, device=self.device)
        else:
            self.netG = SingleSpeedNVP(x_dim, hidden_dim, num_blocks, num_layers, scale=scale, device=self.device)

        if load_model:
            self.path = os.path.join(log_dir, load_model)
            self.netG.load_state_dict(torch.load(os.path.join(self.path, 'models', 'netG.pt'),
                                                 map_location=self.device))
        elif log_dir is not None:
            self.path = log_dir
            for sub in ('models', 'data', 'chains', 'plots'):
                os.makedirs(os.path.join(self.path, sub), exist_ok=True)
        else:
            self.path = None

        # capturable: the optimizer step lives inside the CUDA graph of one training iteration (_GraphedStep)
        self.optimizer = torch.optim.Adam(self.netG.parameters(), lr=learning_rate, weight_decay=weight_decay,
                                          capturable=True)
        self._graphed = {}
        self.fit_log = []
        # fused fitting kernel (nnb_train_epoch): flat parameter vector + Adam moments in state_dict order
        self._arch = (x_dim, hidden_dim, num_layers, num_blocks)
        self._fused = self.flow_kind == 'nvp' and scale == '' and self.engine.train_supported(*self._arch) \
            and os.environ.get('NNB_TRAIN_AUTOGRAD', '0') != '1'
        self._lr, self._wd = learning_rate, weight_decay
        self._adam_m = self._adam_v = None
        self._adam_step = 0
        self._train_seed = None
        self.logger = create_logger(__name__, level=log_level)
        self.log = log
        self.writer = None
        if self.path is not None:
            from torch.utils.tensorboard import SummaryWriter
            self.logger.info(self.netG)
            self.writer = SummaryWriter(self.path)
        self.logger.info('Number of network params: [%s]' % sum(p.numel() for p in self.netG.parameters()))
        self.logger.info('Device [%s]' % self.device)
        self._sync_device()

    # ------------------------------------------------------------------------------------------
    def _sync_device(self, broadcast=True):
        """Export the current weights to the CUDA kernels (and to every rank, rank 0's weights win)."""
        if broadcast:
            dist.broadcast_parameters(self.netG, src=0)       # NCCL over NVLink, one flat buffer
        if self.flow_kind == 'spline':
            if broadcast and dist.is_distributed():
                for m in self.netG.flow.flows:               # the fixed permutations are not parameters
                    if hasattr(m, 'P'):
                        torch.distributed.broadcast(m.P, src=0)
            g = self.netG
            self.engine.set_flow_spline(g.packed_for_kernel(), self.x_dim, g.num_hidden, g.num_blocks, g.num_bins,
                                        g.tail_bound)
            return
        if self.scale == '':
            # the parameters in registration order ARE the nnb_set_flow layout (tests/test_host_logic.py): one device->host
            # copy of the flat vector instead of one per weight / bias tensor
            flat = torch.nn.utils.parameters_to_vector([p.detach() for p in self.netG.parameters()])
            self.engine.set_flow(flat.cpu().numpy(), *self._arch)
        else:
            self.engine.set_flow_from_state_dict(self.netG.state_dict(), scale=self.scale)

    def load_state_dict(self, sd, permutations=None):
        """netG.load_state_dict + export to the kernels.  flow='spline': `permutations` = the fixed P matrices of the 1x1
        convolutions (the reference does not keep them in the state_dict, networks.py:631); loaded weights count as
        initialised (no data-dependent ActNorm initialisation afterwards)."""
        self.netG.load_state_dict(sd)
        if self.flow_kind == 'spline':
            if permutations is not None:
                self.netG.set_permutations(permutations)
            self.netG.mark_initialised()
        self._sync_device()

    # ------------------------------------------------------------------------------------------
    def train(self,
              samples,
              max_iters=10000,
              log_interval=100,
              save_interval=100,
              jitter=0.0,
              validation_fraction=0.1,
              patience=50,
              l2_norm=0.0,
              device_samples=None):
        """Trainer.train of the reference (trainer.py:134-245).  device_samples (not in the reference): the same rows as
        `samples`, float64, already on the device -- NestedSampler keeps its live set there -- which saves the upload."""
        try:
            return self._fit(samples, max_iters, log_interval, save_interval, jitter, validation_fraction, patience,
                             l2_norm, device_samples)
        finally:
            # data/originals.npy is written from `samples` itself by a background thread: it is complete (and `samples` free
            # to change) when train() returns, however it returns
            writer = getattr(self, '_originals_writer', None)
            if writer is not None:
                writer.join()
                self._originals_writer = None

    def _fit(self, samples, max_iters, log_interval, save_interval, jitter, validation_fraction, patience, l2_norm,
             device_samples):
        start_time = time.time()
        samples = np.asarray(samples)

        if self.path:
            self._save_originals(samples)
        # one upload of the (float64) samples serves the jitter search and the train / validation split
        if device_samples is not None:
            assert device_samples.dtype == torch.float64 and tuple(device_samples.shape) == tuple(samples.shape)
            x64 = device_samples
        else:
            x64 = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.float64)).to(self.device)

        if jitter < 0:
            # trainer.py:147-150: 0.2 x the mean of the two nearest "neighbour" distances of every sample, the first
            # being the sample itself (distance 0).  The reference builds a k-d tree on the host, which degenerates to
            # brute force in more than a few dimensions; the same quantity is computed exactly (float64, direct
            # differences) on the device in row blocks.
            training_jitter = .2 * self._mean_two_nearest(samples, x64)
        else:
            training_jitter = jitter

        if self.log:
            self.logger.info('Number of training samples [%d]' % samples.shape[0])
            self.logger.info('Training jitter [%5.4f]' % training_jitter)

        n = samples.shape[0]
        n_valid = int(math.ceil(validation_fraction * n))      # sklearn train_test_split rounding
        perm = np.random.permutation(n)
        perm_d = torch.from_numpy(perm).to(self.device)
        x_valid = x64.index_select(0, perm_d[:n_valid]).float()      # samples[perm[:n_valid]].astype(float32)
        x_train = x64.index_select(0, perm_d[n_valid:]).float()
        del x64

        best_validation_loss = float('inf')
        best_validation_epoch = 0
        params = list(self.netG.parameters())
        best_state = torch.nn.utils.parameters_to_vector(params).detach().clone()   # one flat copy, not 36 tensors
        counter = 0

        fused = self._fused
        if fused:
            flat = torch.nn.utils.parameters_to_vector(params).detach().clone().contiguous()
            if self._adam_m is None:
                self._adam_m, self._adam_v = torch.zeros_like(flat), torch.zeros_like(flat)
                self._train_seed = int(torch.randint(0, 2 ** 62, (1,)).item())    # follows torch.manual_seed
            best_state = flat.clone()
            x_train, x_valid = x_train.contiguous(), x_valid.contiguous()
            # Epochs are queued ahead of the host (nnb_train_epoch_begin / _end): whenever the outcome of epoch e cannot end
            # the fit (patience cannot run out at e, e is not the last epoch) epoch e + 1 is queued BEFORE the losses of e
            # are read, so the device does not idle while the host compares losses, logs and draws the next permutation.
            # `snap[e % 2]` = the weights after epoch e (a stream-ordered copy queued right behind the epoch): what the
            # sequential loop would see in `flat` while it handles epoch e.
            self.engine.train_epoch_drain()
            snap = (torch.empty_like(flat), torch.empty_like(flat))
            queued = 0                          # last epoch of this fit that has been queued
            total0 = self.total_iters

            perms = {'first': 1, 'block': None}

            def perm_of(e):
                """DataLoader(shuffle=True): a fresh uniformly random order of the training rows for every epoch.  The
                orders of up to 64 epochs come from ONE device sort of (epoch << 48 | 48 random bits) keys: torch.randperm
                is a sort per call (and a host round trip below 30 000 rows), which at 50 epochs x 59 000 rows per fit
                was several milliseconds of device time between the epoch kernels."""
                n = x_train.shape[0]
                if n == 0:
                    return None
                if perms['block'] is None or e - perms['first'] >= perms['block'].shape[0]:
                    count = max(1, min(64, max_iters - e + 1, (1 << 25) // n))
                    keys = torch.randint(0, 1 << 48, (count, n), device=x_train.device, dtype=torch.int64)
                    keys += (torch.arange(count, device=x_train.device, dtype=torch.int64) << 48)[:, None]
                    order = torch.argsort(keys.view(-1))               # segment e of the result = flat indices of epoch e
                    perms['block'] = (order % n).view(count, n)
                    perms['first'] = e
                return perms['block'][e - perms['first']]

            def begin(e):
                self._fused_begin(flat, x_train, x_valid, training_jitter, l2_norm, total0 + e, perm_of(e))
                snap[e % 2].copy_(flat)

        for epoch in range(1, max_iters + 1):
            self.total_iters += 1
            if fused:
                if queued < epoch:
                    begin(epoch)
                    queued = epoch
                if self._lookahead and epoch < max_iters and counter + 1 <= patience:
                    begin(epoch + 1)
                    queued = epoch + 1
                train_loss, validation_loss = self._fused_end(x_train.shape[0], x_valid.shape[0])
                flat_now = snap[epoch % 2]
                if not (math.isfinite(train_loss) and math.isfinite(validation_loss)):
                    # A diverged step (inf / NaN loss) would poison the Adam moments for the rest of the run -- every later
                    # retrain would return the old weights.  Go back to the best weights seen and restart the moments.
                    # (The reference has no such guard; with it a rare divergence costs one epoch instead of the run.)
                    # An epoch already queued behind the diverged one started from the poisoned weights: its result is
                    # dropped and the epoch is queued again after the restore.
                    self.logger.warning('Epoch [%i] non-finite loss: restoring the best weights, resetting Adam' % epoch)
                    if queued > epoch:
                        self.engine.train_epoch_end()
                        queued = epoch
                    flat.copy_(best_state)
                    self._adam_m.zero_()
                    self._adam_v.zero_()
                    self._adam_step = 0
                    validation_loss = float('inf')
                    flat_now = flat
            else:
                train_loss = self._train(epoch, x_train, jitter=training_jitter, l2_norm=l2_norm)
                validation_loss = self._validate(epoch, x_valid)

            if validation_loss < best_validation_loss:
                best_validation_epoch = epoch
                best_validation_loss = validation_loss
                if fused:
                    best_state.copy_(flat_now)
                else:
                    best_state = torch.nn.utils.parameters_to_vector(params).detach().clone()
                counter = 0

            if epoch == 1 or epoch % log_interval == 0:
                self.logger.info('Epoch [%i] train loss [%5.4f] validation loss [%5.4f]' % (
                    epoch, train_loss, validation_loss))

            if self.path:
                self.writer.add_scalar('loss', validation_loss, self.total_iters)
                if epoch % save_interval == 0:
                    if fused:
                        _copy_into_params(params, flat_now)
                    torch.save(self.netG.state_dict(), os.path.join(self.path, 'models', 'netG.pt'))

            counter += 1
            if counter > patience:
                self.logger.info('Epoch [%i] ran out of patience' % (epoch))
                if self.path:
                    if fused:
                        _copy_into_params(params, flat_now)
                    torch.save(self.netG.state_dict(), os.path.join(self.path, 'models', 'netG.pt'))
                break

        self.logger.info('Best epoch [%i] validation loss [%5.4f] train time (s) [%5.4f]]'
                         % (best_validation_epoch, best_validation_loss, time.time() - start_time))
        self.best_validation_epoch = best_validation_epoch
        self.best_validation_loss = best_validation_loss
        # run diagnostics (not in the reference): one record per fit
        self.fit_log.append((self.total_iters, int(samples.shape[0]), float(training_jitter), int(best_validation_epoch),
                             float(best_validation_loss)))
        _copy_into_params(params, best_state)
        self._sync_device()

    _lookahead = True      # queue epoch e + 1 before the losses of epoch e are read (False: one epoch at a time; same results)

    def _fused_begin(self, flat, x_train, x_valid, jitter, l2_norm, epoch_id, perm):
        """Trainer._train + Trainer._validate (trainer.py:384-418) as one launch of the fused fitting kernel, queued without
        waiting (nnb_train_epoch_begin).  The l2 penalty's gradient 2 * l2_norm * w is folded into the weight-decay term
        (identical update; the reported loss excludes the penalty in the reference too, trainer.py:396-397)."""
        n, n_valid = x_train.shape[0], x_valid.shape[0]
        self.engine.train_epoch_begin(
            self._arch, flat, self._adam_m, self._adam_v, self._adam_step, x_train if n else None,
            x_valid if n_valid else None, self.batch_size, perm=perm, jitter=jitter, lr=self._lr,
            weight_decay=self._wd + 2.0 * l2_norm, seed=self._train_seed, epoch=epoch_id)
        self._adam_step += (n + self.batch_size - 1) // self.batch_size

    def _fused_end(self, n, n_valid):
        """(train loss, validation loss) of the oldest queued epoch, normalised as the reference does."""
        tl, vs, _ = self.engine.train_epoch_end()
        train_loss = tl / n if n else 0.0
        validation_loss = (vs / n_valid) / n_valid if n_valid else 0.0      # trainer.py:414-418
        return train_loss, validation_loss

    def _mean_two_nearest(self, samples, x64=None):
        """np.mean(cKDTree(samples).query(samples, 2)[0]): mean over samples of (0 + nearest other sample) / 2."""
        if x64 is None:
            x64 = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.float64)).to(self.device)
        return 0.5 * self.engine.mean_nn_distance(x64)

    def _save_originals(self, samples):
        """data/originals.npy (trainer.py:160-161), written by a background thread: at config-4 size it is 16 MB per fit and
        272 fits per run.  The previous write is joined first, so the file on disk is always a complete array."""
        import threading
        prev = getattr(self, '_originals_writer', None)
        if prev is not None:
            prev.join()
        # no private copy: train() joins the writer before it returns, and nobody touches `samples` in between
        path = os.path.join(self.path, 'data', 'originals.npy')
        self._originals_writer = threading.Thread(target=np.save, args=(path, samples), daemon=False)
        self._originals_writer.start()

    def _train(self, epoch, x_train, jitter=0.0, l2_norm=0.0):
        """One epoch (trainer.py:384-403): shuffled mini-batches, jittered inputs, Adam on -mean(log p).  Full batches
        replay a captured CUDA graph (forward + backward + Adam in one launch); a trailing partial batch runs eagerly."""
        self.netG.train()
        n = x_train.shape[0]
        order = torch.randperm(n, device=x_train.device)
        bs = self.batch_size
        if n and getattr(self.netG, 'needs_data_init', lambda: False)():
            # ActNorm's data-dependent initialisation (networks.py:687-693) happens on the first batch the flow sees; done
            # eagerly here because the training step below is replayed from a CUDA graph
            with torch.no_grad():
                first = x_train[order[:bs]]
                self.netG.forward(first + jitter * torch.randn_like(first))
        total = torch.zeros((), device=x_train.device)
        nfull = n // bs
        if nfull and not l2_norm:
            step = self._graphed.get(bs)
            if step is None:
                step = self._graphed[bs] = _GraphedStep(self.netG, self.optimizer, bs, self.x_dim, x_train.device)
            for b in range(nfull):
                total += step(x_train[order[b * bs:(b + 1) * bs]], jitter)
            rest = range(nfull * bs, n, bs)
        else:
            rest = range(0, n, bs)
        for s in rest:
            data = x_train[order[s:s + bs]]
            data = data + jitter * torch.randn_like(data)
            self.optimizer.zero_grad(set_to_none=False)
            loss = -self.netG.log_probs(data).mean()
            total += loss.detach()
            if l2_norm:
                loss = loss + l2_norm * sum((p ** 2).sum() for p in self.netG.parameters())
            loss.backward()
            self.optimizer.step()
        return total.item() / n      # the reference divides by the dataset size (trainer.py:403)

    def _validate(self, epoch, x_valid):
        self.netG.eval()
        if x_valid.shape[0] == 0:
            return 0.0
        with torch.no_grad():
            val = -self.netG.log_probs(x_valid).mean().item()
        return val / x_valid.shape[0]     # trainer.py:418

    # ------------------------------------------------------------------------------------------
    def _as_device(self, a):
        if isinstance(a, np.ndarray):
            return torch.from_numpy(a).float().to(self.device)
        return a.to(self.device, dtype=torch.float32)

    def forward(self, x, to_numpy=False):
        z, log_det_J = self.engine.flow_forward(self._as_device(x))
        if to_numpy:
            return z.cpu().numpy(), log_det_J.cpu().numpy()
        return z, log_det_J

    def inverse(self, z, to_numpy=False):
        z = self._as_device(z)
        if self.flow_kind == 'spline' and z.shape[0] == 1 and bool(self.engine.flow_empty_halves(z).all()):
            # RQS on an empty selection (networks.py:464-465): a coupling transform finds no coordinate of the batch inside
            # the tail bound.  Checked for one-sample batches, the only case in which it happens in practice.
            raise ValueError('No input values')
        x, log_det_J = self.engine.flow_inverse(z)
        if to_numpy:
            return x.cpu().numpy(), log_det_J.cpu().numpy()
        return x, log_det_J

    def get_prior_samples(self, num_samples, to_numpy=False):
        z = torch.randn((num_samples, self.x_dim), device=self.device)
        return z.cpu().numpy() if to_numpy else z

    def get_latent_samples(self, x, to_numpy=False):
        return self.forward(x, to_numpy=to_numpy)[0]

    def get_samples(self, z, to_numpy=False):
        return self.inverse(z, to_numpy=to_numpy)[0]

    def get_synthetic_samples(self, num_samples, to_numpy=False):
        return self.get_samples(self.get_prior_samples(num_samples), to_numpy=to_numpy)

    def log_probs(self, x, to_numpy=False):
        z, log_det_J = self.forward(x)
        lp = -0.5 * (z * z).sum(-1) - 0.5 * self.x_dim * math.log(2 * math.pi) + log_det_J
        return lp.cpu().numpy() if to_numpy else lp

    def plot_samples(self, samples, outfile=None, plot_synthetic=True):
        """Plotting is outside the accelerated path (matplotlib is optional)."""
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            return
