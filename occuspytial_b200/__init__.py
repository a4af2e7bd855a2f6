"""occuspytial_b200 -- B200-native engine behind OccuSpytial's sampler API.

``LogitICARGibbs`` (alias ``PGGibbs``), ``LogitRSRGibbs`` and ``ProbitRSRGibbs`` keep the reference's
constructor and ``sample`` signatures (occuspytial/__init__.py:1-3) and return the same
``PosteriorParameter`` / ``Chain`` objects; the Gibbs sweep runs as hand-written sm_100a CUDA kernels
for thousands of lock-step chains (occuspytial_b200/csrc) behind a C ABI (include/occu_b200.h).
"""
from .gibbs import LogitICARGibbs, LogitRSRGibbs, ProbitRSRGibbs, PGGibbs
from .data import Data

__version__ = '0.1.0'
__all__ = ('LogitICARGibbs', 'LogitRSRGibbs', 'ProbitRSRGibbs', 'PGGibbs', 'Data')
