"""Per-chain sample storage with the reference's ``Chain`` API (occuspytial/chain.py:4-117).

On the engine side draws are recorded on the device for all chains at once and flushed in blocks
(``gibbs/base.py``); a ``Chain`` handed back to the user is then a zero-copy view of one chain's
slice of that block (``Chain.from_arrays``).  The stand-alone append/expand behaviour of the
reference (tested by occuspytial/tests/test_chain.py) is kept.
"""
import numpy as np


class Chain:
    """Container to store parameter chains during sampling.

    Parameters
    ----------
    params : Dict[str, int]
        Parameter name -> number of components.
    size : int
        Capacity (number of draws) of the chain.
    """

    def __init__(self, params, size):
        self.size = int(size)
        self._names = tuple(params)
        self._index = 0
        self._store = {
            name: np.zeros((self.size, width) if width > 1 else self.size)
            for name, width in params.items()
        }

    @classmethod
    def from_arrays(cls, arrays):
        """Wrap already-filled per-parameter arrays (draws along axis 0) without copying."""
        self = cls.__new__(cls)
        self._names = tuple(arrays)
        self._store = dict(arrays)
        first = next(iter(arrays.values()))
        self.size = self._index = int(first.shape[0])
        return self

    @property
    def full(self):
        """All parameters side by side, one row per stored draw (chain.py:33-54)."""
        cols = [a if a.ndim > 1 else a[:, None] for a in self._store.values()]
        return np.concatenate(cols, axis=1)[:self._index]

    def append(self, params):
        """Store one draw of every parameter (chain.py:56-74)."""
        if self._index >= self.size:
            raise ValueError('Chain is full, cannot append any new values')
        for name, value in params.items():
            self._store[name][self._index] = value
        self._index += 1

    def expand(self, size):
        """Grow the capacity by ``size`` draws (chain.py:76-92)."""
        for name, arr in self._store.items():
            pad = np.zeros((size,) + arr.shape[1:])
            self._store[name] = np.concatenate([arr, pad], axis=0)
        self.size += size

    def __getitem__(self, name):
        return self._store[name][:self._index]

    def __len__(self):
        return self._index

    def __repr__(self):
        return f'Chain(params: {self._names}, size: {self._index})'


class ChainViews:
    """Read-only sequence of the per-chain :class:`Chain` objects of a lock-step run.

    The engine returns all chains stacked, ``{name: (chains, draws[, dim])}``; the reference's API is a list of
    ``Chain`` objects (parallel.py:38-42).  Building thousands of views eagerly costs milliseconds per ``sample``
    call, so a view is made (zero-copy) when a chain is first asked for.
    """

    def __init__(self, stacked):
        self._stacked = stacked
        self._n = int(next(iter(stacked.values())).shape[0])
        self._cache = {}

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError('chain index out of range')
        if i not in self._cache:
            self._cache[i] = Chain.from_arrays({k: v[i] for k, v in self._stacked.items()})
        return self._cache[i]

    def __iter__(self):
        return (self[i] for i in range(self._n))

    def __repr__(self):
        return f'ChainViews({self._n} chains)'
