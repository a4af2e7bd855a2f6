"""Stand-in for polyagamma==1.2.0 (absent, no network) -- TEST INFRASTRUCTURE.

Only the call shape the reference uses is supported (logit.py:191-193, 202-204):
``random_polyagamma(1, z, disable_checks=True, random_state=Generator)``.
Draws come from the C restatement oracle/pg_devroye.c, seeded with one 64-bit
integer pulled from the caller's Generator (so same Generator state => same draws).
"""
import ctypes
import os

import numpy as np

_lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'liboccu_oracle.so'))
_lib.occu_oracle_pg1_fill.argtypes = [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
_lib.occu_oracle_pg1_fill.restype = None


def random_polyagamma(h=1, z=0, *, size=None, out=None, method=None, disable_checks=False, random_state=None):
    if h != 1 or size is not None or out is not None:
        raise NotImplementedError('shim supports random_polyagamma(1, z, ...) only')
    rng = np.random.default_rng(random_state)
    zz = np.ascontiguousarray(z, dtype=np.float64)
    res = np.empty_like(zz)
    seed = int(rng.integers(0, 2**63 - 1))
    _lib.occu_oracle_pg1_fill(seed, zz.size, zz.ctypes.data, res.ctypes.data)
    return res if zz.ndim else float(res)
