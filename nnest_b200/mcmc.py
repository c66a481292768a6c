"""MCMCSampler with the reference's interface (nnest/mcmc.py:18-126): fit the flow to standardised training
samples, then run Metropolis-Hastings chains in its latent space -- all chain steps inside the fused CUDA
kernel (NNB_MODE_MH).  The reference derives MCMCSampler from its emcee-based EnsembleSampler; that class
(third-party moves, work in progress upstream) is outside the accelerated path, so this one derives from
Sampler directly.
"""
from __future__ import division, print_function

import logging

import numpy as np

from .sampler import Sampler


class MCMCSampler(Sampler):

    def __init__(self,
                 x_dim,
                 loglike,
                 prior=None,
                 append_run_num=True,
                 hidden_dim=16,
                 num_slow=0,
                 num_derived=0,
                 batch_size=100,
                 flow='spline',
                 num_blocks=3,
                 num_layers=1,
                 learning_rate=0.001,
                 log_dir='logs/test',
                 base_dist=None,
                 scale='',
                 use_gpu=True,
                 trainer=None,
                 transform_prior=True,
                 oversample_rate=-1,
                 log_level=logging.INFO,
                 param_names=None,
                 seed=0):
        super(MCMCSampler, self).__init__(x_dim, loglike, append_run_num=append_run_num,
                                          hidden_dim=hidden_dim, num_slow=num_slow,
                                          num_derived=num_derived, batch_size=batch_size, flow=flow,
                                          num_blocks=num_blocks, num_layers=num_layers, learning_rate=learning_rate,
                                          log_dir=log_dir, use_gpu=use_gpu, base_dist=base_dist, scale=scale,
                                          trainer=trainer, prior=prior, transform_prior=transform_prior,
                                          log_level=log_level, oversample_rate=oversample_rate,
                                          param_names=param_names, seed=seed)
        self.sampler = 'mcmc'

    def run(
            self,
            mcmc_steps,
            mcmc_num_chains,
            training_samples,
            mcmc_dynamic_step_size=True,
            stats_interval=100,
            output_interval=None,
            initial_jitter=0.01,
            final_jitter=0.01,
            init_samples=None,
            train_iters=10000,
            thin=1):
        mean = np.mean(training_samples, axis=0)
        std = np.std(training_samples, axis=0)
        training_samples = (training_samples - mean) / std          # mcmc.py:107-110
        self.transform = lambda x: x * std + mean                   # mcmc.py:111 (float64: promotes the likelihood)
        self.trainer.train(training_samples, max_iters=train_iters, jitter=initial_jitter)

        # as the reference (mcmc.py:114-116) the step size stays fixed at 2/sqrt(d): dynamic_step_size is not passed
        # `samples * std + mean` (mcmc.py:117, float64) is evaluated on the device while the trace is staged to the host;
        # thin=k (extension, default 1 = the reference's full trace) keeps every k-th state
        samples, latent_samples, derived_samples, loglikes, scale, ncall = self._mcmc_sample(
            mcmc_steps, num_chains=mcmc_num_chains, stats_interval=stats_interval, output_interval=output_interval,
            init_samples=init_samples, thin=thin, sample_affine=(std, mean))
        if mcmc_steps > 1:
            # mcmc.py:119-120; the statistics of the transformed trace are taken on the device copy
            self._chain_stats(None, trace=self._device_trace, t_scale=std, t_shift=mean)
        self._device_trace = None

        self.samples = np.concatenate((samples, derived_samples), axis=2) if derived_samples.shape[2] else samples
        self.latent_samples = latent_samples
        self.loglikes = loglikes
        self.logger.info("ncall: {:d}\n".format(int(self.total_calls)))
