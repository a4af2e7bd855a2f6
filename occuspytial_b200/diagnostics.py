"""Convergence diagnostics from per-chain summaries (host side, NumPy).

The device computes, per chain and component, the block ``{mean, M2, mean_h1, M2_h1, mean_h2, M2_h2,
acov_0..acov_L}`` (occu_chain_summaries, include/occu_b200.h); ranks all-gather those blocks
(``gibbs/parallel.py``) and the functions below turn them into split-R-hat and ESS without ever
moving the draws.  ``summary_frame`` is the arviz-free stand-in used by ``PosteriorParameter.summary``.
"""
import numpy as np


def summaries_from_draws(x, max_lag):
    """Host reference of occu_chain_summaries for x of shape (chains, draws): (chains, 6 + max_lag + 1)."""
    x = np.asarray(x, dtype=np.float64)
    chains, draws = x.shape
    half = draws // 2
    out = np.zeros((chains, 6 + max_lag + 1))
    parts = (x, x[:, :half], x[:, draws - half:])
    for k, part in enumerate(parts):
        m = part.mean(axis=1)
        out[:, 2 * k] = m
        out[:, 2 * k + 1] = ((part - m[:, None]) ** 2).sum(axis=1)
    d = x - out[:, :1]
    for lag in range(max_lag + 1):
        out[:, 6 + lag] = (d[:, :draws - lag] * d[:, lag:]).sum(axis=1) / draws
    return out


def split_rhat(summ, draws):
    """Classic split-R-hat (Gelman et al. 2013) from summary blocks of shape (chains, 6 + L + 1)."""
    half = draws // 2
    means = np.concatenate([summ[:, 2], summ[:, 4]])
    variances = np.concatenate([summ[:, 3], summ[:, 5]]) / (half - 1)
    W = variances.mean()
    B = half * means.var(ddof=1)
    var_plus = (half - 1) / half * W + B / half
    return float(np.sqrt(var_plus / W)) if W > 0 else float('nan')


def ess(summ, draws):
    """Multi-chain ESS with Geyer's initial-monotone-sequence truncation from the autocovariances."""
    chains = summ.shape[0]
    means = summ[:, 0]
    acov = summ[:, 6:]                        # (chains, L + 1), biased (1/draws) estimator
    chain_var = acov[:, 0] * draws / (draws - 1)
    W = chain_var.mean()
    var_plus = W * (draws - 1) / draws
    if chains > 1:
        var_plus += means.var(ddof=1)
    if not var_plus > 0:
        return float('nan')
    rho = 1.0 - (W - acov.mean(axis=0)) / var_plus
    rho[0] = 1.0
    L = rho.size
    tau_hat = -1.0
    prev = np.inf
    t = 0
    while t + 1 < L:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        pair = min(pair, prev)
        tau_hat += 2.0 * pair
        prev = pair
        t += 2
    tau_hat = max(tau_hat, 1.0 / np.log10(max(chains * draws, 10)))
    return float(chains * draws / tau_hat)


def hdi(x, prob=0.94):
    """Highest-density interval of a flat sample."""
    x = np.sort(np.asarray(x).ravel())
    n = x.size
    k = max(1, int(np.floor(prob * n)))
    if k >= n:
        return float(x[0]), float(x[-1])
    widths = x[k:] - x[:n - k]
    i = int(np.argmin(widths))
    return float(x[i]), float(x[i + k])


def pooled_moments(summ, draws):
    """Pooled mean and standard deviation (ddof = 1) of all draws from the per-chain blocks (equal chain lengths)."""
    chains = summ.shape[0]
    mean = summ[:, 0].mean()
    m2 = summ[:, 1].sum() + draws * ((summ[:, 0] - mean) ** 2).sum()
    total = chains * draws
    return float(mean), float(np.sqrt(m2 / (total - 1))) if total > 1 else float('nan')


def summary_frame(data, max_lag=None, summaries=None):
    """pandas.DataFrame with arviz-like columns for ``{name: (chains, draws[, dim])}``.

    ``summaries``: optional per-chain summary blocks ``{name: (chains, dim, 6 + L + 1)}`` (device-computed, possibly
    covering more chains than ``data`` holds locally); mean, sd, mcse_mean, ess_bulk and r_hat are then derived from
    them, only the HDI is taken from the draws in ``data``."""
    import pandas as pd
    rows, index = [], []
    for name, arr in data.items():
        arr = np.asarray(arr, dtype=np.float64)
        if arr.ndim == 2:
            arr = arr[:, :, None]
        chains, draws, dim = arr.shape
        L = min(draws - 1, 250) if max_lag is None else max_lag
        blocks = None if summaries is None or name not in summaries else np.asarray(summaries[name])
        for k in range(dim):
            x = arr[:, :, k]
            if blocks is not None:
                summ = blocks[:, k, :]
                mean, sd = pooled_moments(summ, draws)
            else:
                summ = summaries_from_draws(x, L) if draws > 3 else None
                mean, sd = x.mean(), x.std(ddof=1) if x.size > 1 else np.nan
            lo, hi = hdi(x) if x.size else (np.nan, np.nan)
            if summ is not None:
                n_eff = ess(summ, draws)
                rhat = split_rhat(summ, draws) if draws >= 4 else np.nan
            else:
                n_eff, rhat = np.nan, np.nan
            mcse = sd / np.sqrt(n_eff) if n_eff and n_eff > 0 else np.nan
            rows.append({'mean': mean, 'sd': sd, 'hdi_3%': lo, 'hdi_97%': hi, 'mcse_mean': mcse,
                         'ess_bulk': n_eff, 'r_hat': rhat})
            index.append(f'{name}[{k}]' if dim > 1 else name)
    return pd.DataFrame(rows, index=index)
