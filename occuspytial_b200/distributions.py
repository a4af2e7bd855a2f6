"""Host-callable versions of the reference's native routines (occuspytial/distributions.pyx), executed
by the CUDA engine: ``precision_mvnorm`` (distributions.pyx:42-110) and ``ensure_sums_to_zero``
(distributions.pyx:24-39).  Batched forms take a leading batch axis; the scalar forms keep the
reference's behaviour, including overwriting ``prec`` with its Cholesky factor (tested by
occuspytial/tests/test_distributions.py:17).
"""
import numpy as np

from . import _lib

__all__ = ('precision_mvnorm', 'ensure_sums_to_zero', 'precision_mvnorm_batched')


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.EngineError('occuspytial_b200 needs a CUDA device; there is no CPU fallback')
    return torch


def precision_mvnorm_batched(b, prec, z=None, seed=0):
    """x[i] = prec[i]^-1 b[i] + L[i]^-T z[i]; returns (x, L) with L the lower Cholesky factors.

    b (m, d), prec (m, d, d), z (m, d) injected normals or None (Philox stream ``seed``)."""
    torch = _torch()
    lib = _lib.load()
    b = np.ascontiguousarray(b, dtype=np.float64)
    prec = np.ascontiguousarray(prec, dtype=np.float64)
    m, d = b.shape
    tb = torch.from_numpy(b).cuda()
    tp = torch.from_numpy(prec).cuda()
    tz = None if z is None else torch.from_numpy(np.ascontiguousarray(z, dtype=np.float64)).cuda()
    tx = torch.empty_like(tb)
    stream = torch.cuda.current_stream().cuda_stream
    fails = _lib.check(lib.occu_precision_mvnorm_batched(d, m, tb.data_ptr(), tp.data_ptr(),
                                                         None if tz is None else tz.data_ptr(), int(seed),
                                                         tx.data_ptr(), stream), 'occu_precision_mvnorm_batched')
    if fails:
        raise RuntimeError('Cholesky factorization/solver failed!')
    return tx.cpu().numpy(), np.tril(tp.cpu().numpy())


def precision_mvnorm(b, prec, random_state=None):
    """Draw from N(prec^-1 b, prec^-1); ``prec`` (C-contiguous float64) is overwritten in its lower
    triangle by the Cholesky factor, like the reference."""
    rng = np.random.default_rng(random_state)
    b = np.ascontiguousarray(b, dtype=np.float64)
    if not (isinstance(prec, np.ndarray) and prec.dtype == np.float64 and prec.flags.c_contiguous and prec.ndim == 2):
        raise ValueError('prec must be a C-contiguous 2-d float64 array')
    z = rng.standard_normal(b.shape[0])
    x, L = precision_mvnorm_batched(b[None], prec[None], z[None])
    prec[np.tril_indices_from(prec)] = L[0][np.tril_indices_from(prec)]
    return x[0]


def ensure_sums_to_zero(x, z, out):
    """out = x - (sum(x) / sum(z)) z  (distributions.pyx:24-39), on the device."""
    torch = _torch()
    lib = _lib.load()
    tx = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)[None]).cuda()
    tz = torch.from_numpy(np.ascontiguousarray(z, dtype=np.float64)[None]).cuda()
    to = torch.empty_like(tx)
    _lib.check(lib.occu_ensure_sums_to_zero_batched(tx.shape[1], 1, tx.data_ptr(), tz.data_ptr(), to.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream),
               'occu_ensure_sums_to_zero_batched')
    out[:] = to.cpu().numpy()[0]
    return out
