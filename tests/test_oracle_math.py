"""The oracle's arithmetic, pinned against published known answers and high-precision references (CPU only)."""
import numpy as np
import pytest

mpmath = pytest.importorskip("mpmath")
mpmath.mp.prec = 200

# Random123 (Salmon et al., SC'11) known-answer vectors for philox4x32, 10 rounds: (counter, key) -> output
PHILOX_KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers(oracle):
    for ctr, key, want in PHILOX_KAT:
        assert oracle.philox(ctr, key) == want


def _philox_py(ctr, key):
    """straight from the published round function, independent of the C code"""
    c = list(ctr)
    k = list(key)
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return tuple(c)


def test_philox_matches_python_restatement(oracle):
    rng = np.random.default_rng(7)
    for _ in range(300):
        ctr = tuple(int(x) for x in rng.integers(0, 2**32, 4))
        key = tuple(int(x) for x in rng.integers(0, 2**32, 2))
        assert oracle.philox(ctr, key) == _philox_py(ctr, key)


def test_uniform53_is_open_interval_and_exact(oracle):
    L = oracle.lib()
    assert L.oracle_uniform53(0, 0) == 2.0**-53
    assert L.oracle_uniform53(0xFFFFFFFF, 0xFFFFFFFF) == 1.0 - 2.0**-53
    rng = np.random.default_rng(3)
    for _ in range(1000):
        lo, hi = (int(x) for x in rng.integers(0, 2**32, 2))
        m = ((hi << 32) | lo) >> 12
        assert L.oracle_uniform53(lo, hi) == (2 * m + 1) / 2.0**53      # exact in binary64


def test_neg2log_accuracy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(5)
    """|error| <= 2^-51 |x| + 2e-18: relative accuracy everywhere except in the last table interval below 1, where
    logc + log1p(r) cancels and the error is absolute (x ~ 1e-16 there, i.e. |z| ~ 1e-8: irrelevant to a timer)."""
    us = [2.0**-53, 1 - 2.0**-53, 0.5, 0.6875, 0.68749999, 1 / 3, 1e-9, 0.999, 0.99999999]
    us += [float(L.oracle_uniform53(int(a), int(b))) for a, b in rng.integers(0, 2**32, (4000, 2))]
    for u in us:
        got = mpmath.mpf(L.oracle_neg2log(u))
        want = -2 * mpmath.log(mpmath.mpf(u))
        assert abs(got - want) <= mpmath.mpf(2) ** -51 * want + mpmath.mpf("2e-18"), u


def test_sincos_accuracy_all_sectors(oracle):
    """sin/cos(2 pi v / 2^64): 8 sector bits + 52 offset bits, offset = fma(d, pi/128, -(3/2 - 2^-53) pi/128)"""
    rng = np.random.default_rng(11)
    worst = 0.0
    vs = [int(x) for x in rng.integers(0, 2**64, 4000, dtype=np.uint64)]
    vs += [j << 56 for j in range(256)] + [(j << 56) | ((1 << 56) - 1) for j in range(256)]
    vs += [(j << 56) | (1 << 55) for j in range(0, 256, 7)]
    for v in vs:
        s, c = oracle.sincos2pi(v)
        j, m = v >> 56, (v >> 4) & ((1 << 52) - 1)
        ang = (mpmath.mpf(j) + mpmath.mpf(2 * m + 1) / 2**53) * mpmath.pi / 128      # the angle the bits stand for
        worst = max(worst, float(abs(mpmath.sin(ang) - s)), float(abs(mpmath.cos(ang) - c)))
        assert abs(s * s + c * c - 1.0) < 1e-15
    assert worst < 3e-16


def test_normal_pair_moments(oracle):
    """the two components are standard normal and uncorrelated (200k Philox blocks)"""
    n = 200_000
    z = np.empty((n, 2))
    for i in range(n):
        z[i] = oracle.normal_pair(_philox_py((i, 0, 1, 0), (0x5EED0001, 0)))
    assert abs(z.mean()) < 0.01
    assert abs(z.var() - 1.0) < 0.01
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.01
    from scipy import stats
    assert stats.kstest(z[:, 0], "norm").pvalue > 1e-3
    assert stats.kstest(z[:, 1], "norm").pvalue > 1e-3


def test_ziggurat_law(oracle):
    """Division timers draw their standard normal with the ziggurat method: 4e6 draws against the normal law - moments,
    KS, a chi-square over 200 equiprobable cells plus the two tails beyond r where the tail sampler takes over, sibling
    independence - and the trial statistics the published construction predicts (99.64 % of the trials accepted with 512 layers)."""
    from scipy import stats
    n = 4_000_000
    z, trials = oracle.zig_fill(0x5EED0002, n)
    assert abs(n / trials - 0.996383) < 2e-4                  # sqrt(pi/2) / (512 V)
    assert abs(z.mean()) < 4 / np.sqrt(n) and abs(z.var() - 1.0) < 6 * np.sqrt(2.0 / n)
    assert abs((z**3).mean()) < 5 * np.sqrt(15.0 / n) and abs((z**4).mean() - 3.0) < 5 * np.sqrt(96.0 / n)
    assert stats.kstest(z, "norm").pvalue > 1e-3
    r = 3.8520461503683912
    edges = np.concatenate(([-np.inf, -r], stats.norm.ppf(np.linspace(0, 1, 201)[1:-1]), [r, np.inf]))
    edges = np.unique(edges)
    obs = np.histogram(z, bins=edges)[0]
    exp = n * np.diff(stats.norm.cdf(edges))
    chi2 = ((obs - exp) ** 2 / exp).sum()
    assert stats.chi2.sf(chi2, len(exp) - 1) > 1e-3, (chi2, len(exp))
    tail = np.abs(z) > r
    assert abs(tail.sum() - n * 2 * stats.norm.sf(r)) < 5 * np.sqrt(n * 2 * stats.norm.sf(r))
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 5 / np.sqrt(n / 2)      # daughters 0 and 1 of one block
    assert abs((z < 0).mean() - 0.5) < 4 * 0.5 / np.sqrt(n)
