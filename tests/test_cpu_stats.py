"""CPU: device-side statistics (surrdamh_b200/stats.py, torch on CPU tensors here) against direct numpy restatements of
what the reference computes with emcee / numpy (visualization_and_analysis.py:437,580-594,720-747)."""
import numpy as np
import torch

from surrdamh_b200 import stats


def _acf_numpy(x):                      # emcee.autocorr.function_1d
    n = 1
    while n < len(x):
        n <<= 1
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    return acf / acf[0]


def _auto_window(taus, c):              # visualization_and_analysis.py:720-724
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return np.argmin(m)
    return len(taus) - 1


def test_autocorrelation_matches_reference_formulas():
    rs = np.random.RandomState(0)
    K, C = 3000, 5
    x = np.zeros((K, C))
    for c in range(C):                  # AR(1) chains with different correlation lengths
        rho = 0.5 + 0.09 * c
        e = rs.standard_normal(K)
        for k in range(1, K):
            x[k, c] = rho * x[k - 1, c] + e[k]
    acf = stats.autocorr_function(torch.as_tensor(x))
    tau = stats.autocorr_time(acf, c=5.0)
    for c in range(C):
        ref = _acf_numpy(x[:, c])
        assert np.allclose(acf[:, c].numpy(), ref, rtol=1e-9, atol=1e-11)
        taus = 2.0 * np.cumsum(ref) - 1.0
        assert abs(tau[c].item() - taus[_auto_window(taus, 5.0)]) < 1e-8
        rho = 0.5 + 0.09 * c
        if c < 3:       # statistical sanity on the shorter correlation lengths: AR(1) has tau = (1+rho)/(1-rho)
            assert abs(tau[c].item() - (1 + rho) / (1 - rho)) < 0.4 * (1 + rho) / (1 - rho)


def test_current_trace_and_cpus():
    flags = torch.tensor([[0, 1], [1, 2], [2, 0], [1, 1]], dtype=torch.uint8)           # [K=4][C=2]
    prop = torch.arange(4 * 1 * 2, dtype=torch.float64).reshape(4, 1, 2) + 10.0           # d = 1
    init = torch.tensor([[-1.0, -2.0]])
    cur = stats.current_trace(flags, prop, init)
    assert cur[:, 0, 0].tolist() == [-1.0, 12.0, 12.0, 16.0]
    assert cur[:, 0, 1].tolist() == [11.0, 11.0, 11.0, 17.0]
    counters = np.array([[10, 30, 60, 0], [20, 20, 60, 0]])
    tau = torch.tensor([4.0, 2.0], dtype=torch.float64)
    want = np.mean([(40 / 100 + 0.1) * 4.0, (40 / 100 + 0.1) * 2.0])
    assert abs(stats.cpus(counters, tau, 0.1).item() - want) < 1e-12
