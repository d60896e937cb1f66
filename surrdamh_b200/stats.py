"""Chain statistics on the device, for runs whose chain count makes the reference's CSV + pandas + emcee
post-processing (surrDAMH/modules/visualization_and_analysis.py) impractical.

  chain_moments        weighted mean / covariance of the samples   (what Samples.calculate_properties derives from the
                        decompressed chains, :82-91,738-747) -- from the running sums the chain kernel keeps
  current_trace        the chain's CURRENT sample at every step, rebuilt from the per-step trace (flags + proposals)
  autocorr_function    normalised autocorrelation function per chain, FFT based (emcee.autocorr.function_1d, used at :437)
  autocorr_time        integrated autocorrelation time with Sokal's automatic window (auto_window / autocorr_FM, :720-736)
  cpus                 cost per uncorrelated sample, (full evaluations / all steps + cost ratio) * tau  (:580-594)

torch only (plumbing: FFTs and reductions over tensors that already live in HBM); works on CPU tensors as well.
"""
import torch


def chain_moments(block, steps=None):
    """(mean [d], cov [d x d]) over all chains and all steps accumulated in block.moments."""
    if block.moments is None:
        raise ValueError("the ChainBlock was created without moments=True")
    d = block.d
    n = float(block.C) * float(steps if steps is not None else block.moment_steps)
    mom = block.moments.sum(dim=1) / n
    mean = mom[:d]
    second = torch.zeros((d, d), dtype=mom.dtype, device=mom.device)
    q = d
    for i in range(d):
        for j in range(i, d):
            second[i, j] = second[j, i] = mom[q]
            q += 1
    return mean, second - torch.outer(mean, mean)


def current_trace(flags, prop, initial):
    """flags [K][C] (1 = accepted), prop [K][d][C], initial [d][C] -> the current sample after every step, [K][d][C]."""
    K, C = flags.shape
    acc = flags == 1
    steps = torch.arange(1, K + 1, device=flags.device).reshape(K, 1).expand(K, C)
    last = torch.cummax(torch.where(acc, steps, torch.zeros_like(steps)), dim=0).values      # 0: still the initial sample
    allv = torch.cat([initial.unsqueeze(0), prop], dim=0)                                    # index 0 = initial
    idx = last.unsqueeze(1).expand(K, prop.shape[1], C)
    return torch.gather(allv, 0, idx)


def autocorr_function(x):
    """x [K][C] (time along dim 0) -> normalised autocorrelation function [K][C] of every chain."""
    K = x.shape[0]
    n = 1
    while n < K:
        n <<= 1
    xc = x - x.mean(dim=0, keepdim=True)
    f = torch.fft.fft(xc, n=2 * n, dim=0)
    acf = torch.fft.ifft(f * torch.conj(f), dim=0)[:K].real
    return acf / acf[0:1]


def autocorr_time(acf, c=5.0):
    """acf [K][C] -> tau [C]: 2 cumsum(acf) - 1 at the first lag m with m >= c * tau(m) (the last lag if there is none)."""
    K = acf.shape[0]
    taus = 2.0 * torch.cumsum(acf, dim=0) - 1.0
    lags = torch.arange(K, device=acf.device, dtype=taus.dtype).reshape(K, 1)
    ok = lags < c * taus                                   # still inside the window
    any_false = (~ok).any(dim=0)
    first_false = torch.argmin(ok.to(torch.int8), dim=0)   # first lag where the window condition fails
    window = torch.where(any_false, first_false, torch.full_like(first_false, K - 1))
    return torch.gather(taus, 0, window.unsqueeze(0)).squeeze(0)


def cpus(counters, tau, surr_cost_ratio=0.0):
    """counters [C][>=3] (accepted, rejected, prerejected), tau [C] -> mean cost per uncorrelated sample."""
    c = torch.as_tensor(counters, dtype=torch.float64, device=tau.device)
    full = c[:, 0] + c[:, 1]
    allsteps = c[:, 0] + c[:, 1] + c[:, 2]
    return ((full / allsteps + surr_cost_ratio) * tau).mean()
