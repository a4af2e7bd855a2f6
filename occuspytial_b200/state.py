"""Attribute containers of the samplers (reference: occuspytial/gibbs/state.py:4-27)."""
from types import SimpleNamespace


class _Storage(SimpleNamespace):
    """Namespace whose attributes can also be read (not written) by item access."""

    def __getitem__(self, key):
        return self.__dict__[key]


class State(_Storage):
    """Mutable sampler state; iterating yields the attribute names in insertion order."""

    def __iter__(self):
        return iter(self.__dict__)


class FixedState(_Storage):
    """Write-once container for quantities that stay constant during sampling."""

    def __setattr__(self, name, value):
        if name in self.__dict__:
            raise KeyError('cannot change attributes already set')
        super().__setattr__(name, value)
