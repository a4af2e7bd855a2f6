"""Ragged site-visit data.

``Data`` keeps the reference's dictionary-backed API (occuspytial/data.pyx:34-147: ``surveyed``,
``visits``, ``__getitem__`` with multi-site concatenation, ``len``, pickling) for user code, and
adds ``to_csr`` which flattens the dictionary ONCE into the CSR layout the device kernels read
(rows grouped by site in site-id order, ``visit_ptr`` of length n + 1).  The per-iteration
gather/concatenate of the reference (logit.py:187-189, 221) does not exist in the engine.
"""
import numpy as np


class Data:
    """Container for detection data keyed by site number."""

    def __init__(self, data):
        if not isinstance(data, dict):
            raise TypeError('data must be a dict of site -> array')
        self._data = data
        self.surveyed = list(data)

    def visits(self, sites):
        """Number of visits of one site (int) or of each site in a list / tuple (tuple of ints)."""
        if isinstance(sites, (list, tuple)):
            return tuple(self._data[s].shape[0] for s in sites)
        return self._data[sites].shape[0]

    def __getitem__(self, sites):
        if isinstance(sites, (list, tuple)):
            return np.concatenate([self._data[s] for s in sites], axis=0)
        return self._data[sites]

    def __len__(self):
        return len(self._data)

    def __reduce__(self):
        return self.__class__, (self._data,)

    def to_csr(self, n, dtype=np.float64):
        """Flatten to (rows, visit_ptr): rows are the per-site arrays stacked in site-id order."""
        sites = np.fromiter(self._data.keys(), dtype=np.int64, count=len(self._data))
        if sites.size and (sites.min() < 0 or sites.max() >= n):
            raise ValueError('site number out of range')
        order = np.argsort(sites, kind='stable')
        counts = np.zeros(n, dtype=np.int64)
        arrays = [np.asarray(self._data[int(sites[k])]) for k in order]
        counts[sites[order]] = [a.shape[0] for a in arrays]
        visit_ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(counts, out=visit_ptr[1:])
        if arrays:
            rows = np.ascontiguousarray(np.concatenate(arrays, axis=0), dtype=dtype)
        else:
            rows = np.zeros((0,), dtype=dtype)
        return rows, visit_ptr
