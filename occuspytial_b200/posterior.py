"""Posterior container with the reference's API (occuspytial/posterior.py:30-104).

The reference wraps arviz; north_star keeps that behaviour unchanged.  arviz is an optional
dependency here: when it is importable the reference's code path is used verbatim
(``az.convert_to_inference_data(...).posterior``, ``az.summary``, the five plot pass-throughs);
when it is not, the samples are held in a plain mapping and ``summary`` is computed by
``occuspytial_b200.diagnostics`` (split R-hat, ESS, MCSE, HDI) so that ``samples['alpha']`` and
``samples.summary`` keep working.
"""
import numpy as np

try:  # pragma: no cover - depends on the environment
    import arviz as az
except Exception:  # arviz / xarray / matplotlib absent
    az = None


class _Var:
    __slots__ = ('data',)

    def __init__(self, data):
        self.data = data


class PosteriorParameter:
    """Container to store posterior samples, produce plots and summaries.

    Parameters
    ----------
    *chains
        instances of :class:`~occuspytial_b200.chain.Chain`.
    """

    def __init__(self, *chains):
        self.data = self._create_inference_data(chains)
        self.summaries = None

    @classmethod
    def from_arrays(cls, arrays, summaries=None):
        """Build from already stacked arrays ``{name: (chains, draws[, dim])}`` (no per-chain objects).

        ``summaries``: per-chain summary blocks ``{name: (all chains, dim, 6 + L + 1)}`` computed on the device(s)
        where the draws were recorded and, in a multi-process run, all-gathered over the ranks
        (include/occu_b200.h: occu_chain_summaries; gibbs/parallel.py: allgather_summaries).  ``summary`` then takes
        mean, sd, MCSE, ESS and split R-hat from them -- over ALL chains of the job, also those whose draws live in
        another process -- without touching the draws."""
        self = cls.__new__(cls)
        self.data = self._wrap(dict(arrays))
        self.summaries = summaries
        return self

    @staticmethod
    def _wrap(data):
        if az is not None:
            return az.convert_to_inference_data(data).posterior
        return {name: _Var(np.asarray(value)) for name, value in data.items()}

    def _create_inference_data(self, chains):
        names = chains[0]._names
        if len(chains) > 1:
            data = {name: np.stack([c[name] for c in chains]) for name in names}
        else:
            data = {name: chains[0][name][None] for name in names}
        return self._wrap(data)

    def _need_arviz(self, what):
        if az is None:
            raise ImportError(f'{what} needs the optional dependency arviz')

    @property
    def summary(self):
        """Summary statistics: mean, sd, hdi_3%, hdi_97%, mcse_mean, mcse_sd, ess_bulk, ess_tail, r_hat."""
        if az is not None:
            return az.summary(self.data)
        from .diagnostics import summary_frame
        return summary_frame({name: var.data for name, var in self.data.items()}, summaries=self.summaries)

    def plot_trace(self, **kwargs):
        """Plot density and samples values of parameters (arviz.plot_trace)."""
        self._need_arviz('plot_trace')
        return az.plot_trace(self.data, **kwargs)

    def plot_auto_corr(self, **kwargs):
        """Plot the autocorrelation function of each posterior parameter (arviz.plot_autocorr)."""
        self._need_arviz('plot_auto_corr')
        return az.plot_autocorr(self.data, **kwargs)

    def plot_pair(self, **kwargs):
        """Pair plots of the posterior parameters (arviz.plot_pair)."""
        self._need_arviz('plot_pair')
        return az.plot_pair(self.data, **kwargs)

    def plot_density(self, **kwargs):
        """Posterior densities (arviz.plot_posterior)."""
        self._need_arviz('plot_density')
        return az.plot_posterior(self.data, **kwargs)

    def plot_ess(self, **kwargs):
        """Quantile, local or evolution of effective sample size (arviz.plot_ess)."""
        self._need_arviz('plot_ess')
        return az.plot_ess(self.data, **kwargs)

    def __getitem__(self, name):
        return self.data[name].data
