"""oracle/ref_loader.py -- TEST INFRASTRUCTURE, build container only.

Imports the UNMODIFIED reference package from ``/root/reference`` so that the
oracle restatement can be validated against it and golden vectors generated
(tests/golden/make_golden.py).  Nothing on the GPU box may import this module:
``/root/reference`` does not exist there.

What is substituted and why:
  * ``occuspytial.distributions`` / ``occuspytial.data``: the reference's own
    .pyx compiled by oracle/build_ref.py (found through an extra ``__path__``
    entry, no source is copied);
  * ``polyagamma`` (absent third-party dependency): oracle/shims/polyagamma ->
    the C restatement in oracle/pg_devroye.c;
  * ``arviz`` (absent): oracle/shims/arviz, a pass-through container;
  * ``libpysal`` (absent): oracle/shims/libpysal, lattice contiguity restated;
  * ``scipy.linalg.pinvh(cond=...)`` (utils.py:277): the kwarg was renamed
    ``rtol`` in the SciPy installed here; wrapped, same value.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('OCCU_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF, 'occuspytial')) and os.path.isdir(
        os.path.join(HERE, '_ref', 'occuspytial'))


def load_reference():
    """Return the imported reference package ``occuspytial`` (cached in sys.modules)."""
    if 'occuspytial' in sys.modules:
        return sys.modules['occuspytial']
    if not available():
        raise ImportError('reference not available (need /root/reference and `make -C oracle ref`)')
    shims = os.path.join(HERE, 'shims')
    if shims not in sys.path:
        sys.path.insert(0, shims)
    pkg_dir = os.path.join(REF, 'occuspytial')
    spec = importlib.util.spec_from_file_location(
        'occuspytial', os.path.join(pkg_dir, '__init__.py'),
        submodule_search_locations=[pkg_dir, os.path.join(HERE, '_ref', 'occuspytial')])
    mod = importlib.util.module_from_spec(spec)
    sys.modules['occuspytial'] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules['occuspytial']
        raise
    import occuspytial.utils as ref_utils
    from scipy.linalg import pinvh as _pinvh

    def pinvh(a, cond=None, **kw):
        if cond is not None:
            kw['rtol'] = cond
        return _pinvh(a, **kw)

    ref_utils.pinvh = pinvh
    return mod


def run_reference_chains(sampler, size, burnin=0, chains=2, start=None):
    """``sampler.sample(...)`` without the loky process pool (workers cannot import the path-hacked package).

    Mirrors parallel.py:21-41: chain 0 is the sampler itself, chain i>0 a ``copy()`` with a spawned
    seed; each copy is isolated through a pickle round-trip exactly as loky would do.  Returns the
    list of reference ``Chain`` objects.
    """
    import pickle

    samplers = [sampler] + [sampler.copy() for _ in range(chains - 1)]
    out = []
    for s in samplers:
        s = pickle.loads(pickle.dumps(s))
        out.append(s._run(size=size, burnin=burnin, start=start, progressbar=False))
    return out
