"""lat2SW restated from its documented behaviour -- TEST INFRASTRUCTURE."""
import numpy as np
from scipy import sparse


def lat2SW(nrows=3, ncols=5, criterion='rook', row_st=False):
    """Sparse contiguity matrix of an nrows x ncols lattice, site id = i * ncols + j."""
    n = nrows * ncols
    j = np.arange(n) % ncols
    i = np.arange(n) // ncols
    rows, cols = [], []

    def link(mask, off):
        src = np.nonzero(mask)[0]
        rows.append(src)
        cols.append(src + off)

    link(j < ncols - 1, 1)                       # east
    link(i < nrows - 1, ncols)                   # south
    if criterion == 'queen':
        link((i < nrows - 1) & (j < ncols - 1), ncols + 1)   # south-east
        link((i < nrows - 1) & (j > 0), ncols - 1)           # south-west
    elif criterion != 'rook':
        raise ValueError('criterion must be rook or queen')
    r = np.concatenate(rows)
    c = np.concatenate(cols)
    m = sparse.coo_matrix((np.ones(r.size, dtype=np.int64), (r, c)), shape=(n, n)).tocsr()
    m = (m + m.T).tocsr()
    if row_st:
        m = sparse.diags(1.0 / m.sum(axis=1).A1) @ m
    return m
