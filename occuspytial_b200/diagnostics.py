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


REDUCED_HEAD = 5   # leading entries of a reduced block; the mean autocovariances follow


def reduce_blocks(summ):
    """Cross-chain reduction of per-chain summary blocks ``(chains, ..., 6 + L + 1)`` to ``(..., 5 + L + 1)``:
    ``{mean of the chain means, their variance (ddof = 1; 0 for one chain), sum of M2, mean M2 of the 2 x chains
    half-chains, variance (ddof = 1) of the half-chain means, mean autocovariances 0..L}`` -- all that split-R-hat, the
    ESS and the pooled moments need.  Works on a NumPy array or a torch tensor (any device): a multi-GPU run reduces
    the all-gathered blocks where they are and brings back only this."""
    if isinstance(summ, np.ndarray):
        chains = summ.shape[0]
        out = np.empty(summ.shape[1:-1] + (REDUCED_HEAD + summ.shape[-1] - 6,))
        out[..., 0] = summ[..., 0].mean(axis=0)
        out[..., 1] = summ[..., 0].var(axis=0, ddof=1) if chains > 1 else 0.0
        out[..., 2] = summ[..., 1].sum(axis=0)
        out[..., 3] = np.concatenate([summ[..., 3], summ[..., 5]], axis=0).mean(axis=0)
        out[..., 4] = np.concatenate([summ[..., 2], summ[..., 4]], axis=0).var(axis=0, ddof=1)
        out[..., REDUCED_HEAD:] = summ[..., 6:].mean(axis=0)
        return out
    import torch
    chains = summ.shape[0]
    head = [summ[..., 0].mean(dim=0),
            summ[..., 0].var(dim=0, unbiased=True) if chains > 1 else torch.zeros_like(summ[0, ..., 0]),
            summ[..., 1].sum(dim=0),
            torch.cat([summ[..., 3], summ[..., 5]], dim=0).mean(dim=0),
            torch.cat([summ[..., 2], summ[..., 4]], dim=0).var(dim=0, unbiased=True)]
    return torch.cat([torch.stack(head, dim=-1), summ[..., 6:].mean(dim=0)], dim=-1)


def split_rhat_reduced(red, draws):
    """Classic split-R-hat (Gelman et al. 2013) from a reduced block (``reduce_blocks``)."""
    half = draws // 2
    W = red[3] / (half - 1)
    B = half * red[4]
    var_plus = (half - 1) / half * W + B / half
    return float(np.sqrt(var_plus / W)) if W > 0 else float('nan')


def ess_reduced(red, draws, chains):
    """Multi-chain ESS with Geyer's initial-monotone-sequence truncation from the mean autocovariances of a reduced block."""
    acov = np.asarray(red[REDUCED_HEAD:], dtype=np.float64)   # biased (1/draws) estimator, mean over chains
    W = acov[0] * draws / (draws - 1)
    var_plus = W * (draws - 1) / draws
    if chains > 1:
        var_plus += red[1]
    if not var_plus > 0:
        return float('nan')
    rho = 1.0 - (W - acov) / var_plus
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


def pooled_moments_reduced(red, draws, chains):
    """Pooled mean and standard deviation (ddof = 1) of all draws from a reduced block (equal chain lengths)."""
    m2 = red[2] + draws * (chains - 1) * red[1]
    total = chains * draws
    return float(red[0]), float(np.sqrt(m2 / (total - 1))) if total > 1 else float('nan')


def split_rhat(summ, draws):
    """Split-R-hat from per-chain summary blocks of shape (chains, 6 + L + 1)."""
    return split_rhat_reduced(reduce_blocks(np.asarray(summ, dtype=np.float64)), draws)


def ess(summ, draws):
    """Multi-chain ESS from per-chain summary blocks of shape (chains, 6 + L + 1)."""
    summ = np.asarray(summ, dtype=np.float64)
    return ess_reduced(reduce_blocks(summ), draws, summ.shape[0])


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
    summ = np.asarray(summ, dtype=np.float64)
    return pooled_moments_reduced(reduce_blocks(summ), draws, summ.shape[0])


class ChainSummaries:
    """Summary blocks of all chains of a run (``{name: (chains, dim, 6 + L + 1)}``) and their cross-chain reduction.

    The reduction (``reduced[name]``: ``(dim, 5 + L + 1)``, a few hundred numbers) is computed where the blocks are --
    on the device that recorded the draws or, in a multi-process run, on the all-gathered tensor -- and is all that
    ``PosteriorParameter.summary`` needs.  The per-chain blocks themselves stay where they are until somebody indexes
    this object (``summaries['tau']``), which copies them to the host once."""

    def __init__(self, names, dims, chains, packed):
        self.names = list(names)
        self.dims = dict(zip(self.names, dims))
        self.chains = int(chains)
        self._packed = packed            # (chains, sum(dims), 6 + L + 1): NumPy array or torch tensor (any device)
        red = reduce_blocks(packed)
        if not isinstance(red, np.ndarray):
            red = red.cpu().numpy()
        self.reduced = self._split(red, axis=0)
        self._blocks = None

    def _split(self, arr, axis):
        out, o = {}, 0
        for k in self.names:
            d = self.dims[k]
            out[k] = arr[o:o + d] if axis == 0 else arr[:, o:o + d]
            o += d
        return out

    def _load(self):
        if self._blocks is None:
            arr = self._packed if isinstance(self._packed, np.ndarray) else self._packed.cpu().numpy()
            self._blocks = self._split(arr, axis=1)
            self._packed = arr
        return self._blocks

    def __getitem__(self, name):
        return self._load()[name]

    def __contains__(self, name):
        return name in self.dims

    def keys(self):
        return list(self.names)


def summary_frame(data, max_lag=None, summaries=None):
    """pandas.DataFrame with arviz-like columns for ``{name: (chains, draws[, dim])}``.

    ``summaries``: optional :class:`ChainSummaries` (or plain per-chain summary blocks ``{name: (chains, dim, 6 + L + 1)}``),
    device-computed and possibly covering more chains than ``data`` holds locally; mean, sd, mcse_mean, ess_bulk and
    r_hat are then derived from their cross-chain reduction, only the HDI is taken from the draws in ``data``."""
    import pandas as pd
    rows, index = [], []
    for name, arr in data.items():
        arr = np.asarray(arr, dtype=np.float64)
        if arr.ndim == 2:
            arr = arr[:, :, None]
        chains, draws, dim = arr.shape
        L = min(draws - 1, 250) if max_lag is None else max_lag
        reduced = total_chains = None
        if summaries is not None and name in summaries:
            if isinstance(summaries, ChainSummaries):
                reduced, total_chains = summaries.reduced[name], summaries.chains
            else:   # plain per-chain blocks
                blocks = np.asarray(summaries[name], dtype=np.float64)
                reduced, total_chains = reduce_blocks(blocks), blocks.shape[0]
        for k in range(dim):
            x = arr[:, :, k]
            if reduced is not None:
                red, nch = reduced[k], total_chains
                mean, sd = pooled_moments_reduced(red, draws, nch)
            else:
                red = reduce_blocks(summaries_from_draws(x, L)) if draws > 3 else None
                nch = chains
                mean, sd = x.mean(), x.std(ddof=1) if x.size > 1 else np.nan
            lo, hi = hdi(x) if x.size else (np.nan, np.nan)
            if red is not None:
                n_eff = ess_reduced(red, draws, nch)
                rhat = split_rhat_reduced(red, draws) if draws >= 4 else np.nan
            else:
                n_eff, rhat = np.nan, np.nan
            mcse = sd / np.sqrt(n_eff) if n_eff and n_eff > 0 else np.nan
            rows.append({'mean': mean, 'sd': sd, 'hdi_3%': lo, 'hdi_97%': hi, 'mcse_mean': mcse,
                         'ess_bulk': n_eff, 'r_hat': rhat})
            index.append(f'{name}[{k}]' if dim > 1 else name)
    return pd.DataFrame(rows, index=index)
