from .logit import LogitICARGibbs, LogitRSRGibbs, PGGibbs
from .probit import ProbitRSRGibbs

__all__ = ('LogitICARGibbs', 'LogitRSRGibbs', 'ProbitRSRGibbs', 'PGGibbs')
