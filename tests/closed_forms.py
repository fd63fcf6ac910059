"""Known answers that need no oracle: used on the CPU oracle at small depth and on the GPU at depths the oracle cannot
reach in seconds."""
import numpy as np


def check_sigma_zero_tree(simulate, n_seeds: int, k_hi: int):
    """sd = 0: every timer equals the mean, so a seed cell of initial age t0 in (0, mean) gives 2^k leaves at level k
    with k = #{j >= 1 : t0 + j * mean <= t_max}.  With t_max = mean * (k_hi + 1/2) that is k_hi if t0 <= mean / 2, else
    k_hi - 1 - two keys hold everything, each a multiple of its power of two, and the division total follows.
    `simulate(values, freqs, phi, types, t_max)` -> (counts [n_keys][1], divisions, keybase of bin 0)."""
    mean = 10.0
    t_max = mean * k_hi + mean / 2
    values, freqs = np.array([1024.0]), np.array([n_seeds], dtype=np.uint64)
    counts, divisions, keybase = simulate(values, freqs, 1e-9, [[(1.0, mean, 0.0)]], t_max)
    c = counts[:, 0]
    assert set(np.nonzero(c)[0].tolist()) <= {keybase + k_hi - 1, keybase + k_hi}
    a, b = int(c[keybase + k_hi]), int(c[keybase + k_hi - 1])
    assert a % (1 << k_hi) == 0 and b % (1 << (k_hi - 1)) == 0
    na, nb = a >> k_hi, b >> (k_hi - 1)
    assert na + nb == n_seeds
    assert int(divisions) == na * ((1 << k_hi) - 1) + nb * ((1 << (k_hi - 1)) - 1)
    return a, b
