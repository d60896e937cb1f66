"""TEST INFRASTRUCTURE -- numpy restatement of the Gaussian Bayesian problem and
the Gaussian random-walk proposal of the reference.

Follows /root/reference/surrDAMH/modules/classes_SAMPLER.py:
  GaussProblem        <- Problem_Gauss              (:257-316)
  RandomWalkProposal  <- Proposal_GaussRandomWalk   (:222-255)
Operation order is kept (divide by std**2, then multiply, then sum / dot) so the
restatement is bit-identical to the reference on the same machine.
"""
import numpy as np


def _as_vector(x, n):
    # scalar -> filled vector, otherwise taken as given (:260-269, :282-285)
    return np.full((n,), x) if np.isscalar(x) else np.array(x)


class GaussProblem:
    """log posterior = Gaussian log-likelihood + Gaussian log-prior, no constants.

    prior_std / noise_std: scalar or 1-D -> standard deviations; 2-D -> a
    COVARIANCE matrix handed to np.linalg.solve on every call (:300-303, :310-313).
    """

    def __init__(self, no_parameters, prior_mean=0.0, prior_std=1.0, noise_std=1.0,
                 no_observations=None, observations=None, name="default_problem_name"):
        self.no_parameters = no_parameters
        self.prior_mean = _as_vector(prior_mean, no_parameters)
        self.prior_std = _as_vector(prior_std, no_parameters)
        self.prior_is_correlated = self.prior_std.ndim != 1
        self.observations = observations
        if no_observations is None:
            no_observations = len(observations)
        self.no_observations = no_observations
        self.noise_std = _as_vector(noise_std, no_observations)
        self.noise_is_correlated = self.noise_std.ndim != 1
        self.saved_samples_name = name

    def log_likelihood(self, G):
        resid = self.observations - G
        if self.noise_is_correlated:                       # :300-303
            return -0.5 * np.dot(resid, np.linalg.solve(self.noise_std, resid))
        scaled = resid / (self.noise_std ** 2)             # :295-298
        return -0.5 * np.sum(resid * scaled)

    def log_prior(self, sample):
        dev = sample - self.prior_mean
        if self.prior_is_correlated:                       # :310-313
            return -0.5 * np.dot(dev, np.linalg.solve(self.prior_std, dev))
        scaled = dev / (self.prior_std ** 2)               # :305-308
        return -0.5 * np.dot(dev, scaled)

    def log_posterior(self, sample, G):                    # :315-316
        return self.log_likelihood(G) + self.log_prior(sample)


class RandomWalkProposal:
    """u' ~ N(u, diag(std^2)) for scalar / 1-D std, N(u, Sigma) for a 2-D argument.

    `generator` needs .normal(mean, std) (and .multivariate_normal(mean, cov) for
    the 2-D case); the default is numpy's legacy RandomState as in the reference
    (:225).  One generator lives across all stages of a sampler
    (process_SAMPLER.py:46).
    """

    def __init__(self, no_parameters, proposal_std=1.0, seed=0, generator=None):
        self.no_parameters = no_parameters
        self.generator = np.random.RandomState(seed) if generator is None else generator
        self.set_covariance(proposal_std)

    def set_covariance(self, proposal_std=1.0):            # :230-239
        self.proposal_std = _as_vector(proposal_std, self.no_parameters)
        self.is_correlated = self.proposal_std.ndim != 1

    def propose(self, current):
        if self.is_correlated:                             # :253-255
            return self.generator.multivariate_normal(current, self.proposal_std)
        return self.generator.normal(current, self.proposal_std)   # :249-251


class StreamGenerators:
    """Explicit (z, u) streams for one chain: z[step, d], u[step, 2].

    normal() advances the step (one proposal per iteration, :150,:182) and
    rewinds the uniform slot; uniform() hands out slot 0 then slot 1 of the
    current step, so the stage-2 uniform is consumed only when stage 1 passed
    (:193-195) and indexing never drifts between chains.
    """

    def __init__(self, z, u, first_step=0):
        self.z, self.u = z, u
        self.step = first_step - 1
        self.slot = 0

    def normal(self, mean, std):
        self.step += 1
        self.slot = 0
        return mean + std * self.z[self.step]              # loc + scale*gauss, no fma

    def uniform(self, lo=0.0, hi=1.0):
        v = self.u[self.step, self.slot]
        self.slot += 1
        return v
