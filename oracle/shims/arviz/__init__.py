"""Stand-in for arviz (absent) -- TEST INFRASTRUCTURE.  Only what posterior.py:1-61,103 touches."""
import numpy as np


class _Style:
    @staticmethod
    def use(*a, **k):
        return None


style = _Style()


class _Var:
    def __init__(self, data):
        self.data = np.asarray(data)


class _Posterior(dict):
    pass


class _InferenceData:
    def __init__(self, data):
        self.posterior = _Posterior({k: _Var(v) for k, v in data.items()})


def convert_to_inference_data(data, **kwargs):
    return _InferenceData(data)


def summary(data, **kwargs):
    raise NotImplementedError('arviz shim: summary not available')
