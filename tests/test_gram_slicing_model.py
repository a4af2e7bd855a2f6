"""CPU model of the error-free slicing behind the tcgen05 Gram (occuspytial_b200/csrc/gram_tc.cuh).

The CUDA kernel evaluates the reference's FP64 contractions K' diag(omega) K and X' diag(omega) X
(occuspytial/gibbs/logit.py:230, 327-335) by integer slicing.  This test restates the digit decomposition,
the level truncation and the FP64 recombination in NumPy (exact INT64 products stand in for the u8 x s8 -> s32
tensor-core products) and checks the claims the kernel header makes: every digit is in range, no INT32 accumulator
can overflow inside a K chunk, and the result agrees with the FP64 dot product to ~1e-14 of sum |omega Z|.
"""
import numpy as np

SA, SB, LMAX = 6, 6, 7                     # gram_tc.cuh: omega slices, Z slices, kept levels s + t <= LMAX
MAX_SITES_PER_CHUNK = 171 * 64             # gram_tc.cuh: MAX_KB_PER_CHUNK x BK


def column_scale(Z):
    mx = np.abs(Z).max(axis=0)
    f, e = np.frexp(mx)                     # mx = f 2^e, f in [0.5, 1)
    e = e + np.where(f > 0.99, 2, 1)
    return np.where(mx > 0, np.ldexp(1.0, e), 1.0)


def slice_z(Z, c):
    q = np.rint(Z / c * float(1 << (8 * SB))).astype(np.int64)
    digits = []
    for _ in range(SB):                     # balanced base-256 digits, least significant first
        d = ((q + 128) & 0xFF) - 128
        q = (q - d) >> 8
        digits.append(d)
    assert np.all(q == 0)                   # no carry out of the top digit (|Z / c| <= 0.495)
    return digits[::-1]                     # b_1 .. b_SB


def slice_omega(om):
    q = np.minimum(np.floor(om * 2.0 ** 45), 2.0 ** 48 - 1).astype(np.int64)   # (omega / 8) 2^48
    return [(q >> (8 * (SA - 1 - s))) & 0xFF for s in range(SA)]                # a_1 .. a_SA


def test_slicing_reproduces_the_fp64_gram():
    rng = np.random.default_rng(0)
    n, C, M = 6000, 5, 40
    Z = rng.uniform(-2, 2, (n, M)) * np.exp(rng.uniform(-3, 3, M)) * (rng.random((n, M)) > 0.1)
    om = 0.25 * np.exp(3 * (rng.random((n, C)) - 0.7))
    om[rng.random((n, C)) < 1e-3] = 7.9999                                      # near the saturation bound
    c = column_scale(Z)
    b = slice_z(Z, c)
    a = slice_omega(om)
    for d in b:
        assert d.min() >= -128 and d.max() <= 127
    for d in a:
        assert d.min() >= 0 and d.max() <= 255
    assert n <= MAX_SITES_PER_CHUNK
    G = np.zeros((C, M))
    for lev in range(LMAX, 1, -1):                                              # small terms first, as the epilogue
        acc = np.zeros((C, M), dtype=np.int64)
        for s in range(1, SA + 1):
            t = lev - s
            if 1 <= t <= SB:
                acc += a[s - 1].T @ b[t - 1]
        assert np.abs(acc).max() < 2 ** 31                                      # the INT32 accumulator holds it
        G += acc.astype(np.float64) * 2.0 ** (-8 * lev)
    G *= 8.0 * c
    ref = om.T @ Z
    scale = np.abs(om).T @ np.abs(Z)
    assert np.max(np.abs(G - ref) / scale) < 1e-13


def test_chunk_bound_excludes_int32_overflow():
    # the fullest level holds min(SA, SB, LMAX - 1) pairs of |a b| <= 255 x 128 per site
    pairs = min(SA, SB, LMAX - 1)
    assert MAX_SITES_PER_CHUNK * pairs * 255 * 128 < 2 ** 31
