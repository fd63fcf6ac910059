"""Deterministic synthetic inputs for the five BASELINE.json configurations.

The histogram has 1024 log-spaced channels v_i = 10^(4 i / 1023) whose values are round-tripped through
"%.10g" text (what both `procell` binaries read); the N cells are apportioned (largest remainder) over a
Gaussian in channel space, centre 768, width 48 channels, truncated to +-4 widths; every other channel is an
explicit "v 0" line, which exercises the zero-frequency skip of parser.cu:110-111.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

N_CHANNELS = 1024
CENTRE = 768
WIDTH = 48


def synthetic_histogram(n_cells: int, centre: int = CENTRE, width: int = WIDTH, n_channels: int = N_CHANNELS):
    """Returns (values float64[n_channels], freqs uint64[n_channels]); sum(freqs) == n_cells."""
    values = np.array([float("%.10g" % (10.0 ** (4.0 * i / (n_channels - 1)))) for i in range(n_channels)])
    idx = np.arange(n_channels)
    w = np.exp(-0.5 * ((idx - centre) / float(width)) ** 2)
    w[np.abs(idx - centre) > 4 * width] = 0.0
    share = w / w.sum() * n_cells
    base = np.floor(share).astype(np.int64)
    rem = int(n_cells - base.sum())
    if rem > 0:
        order = np.argsort(-(share - base), kind="stable")
        base[order[:rem]] += 1
    assert int(base.sum()) == n_cells
    return values, base.astype(np.uint64)


def histogram_text(values, freqs) -> str:
    return "".join("%.10g %d\n" % (v, int(f)) for v, f in zip(values, freqs))


def types_text(types) -> str:
    return "".join("%.10g %.10g %.10g\n" % tuple(t) for t in types)


TYPES_CONFIG1 = [(0.53, 48.33, 21.6), (0.29, 86.3, 26.8), (0.18, -1.0, -1.0)]
TYPES_CONFIG2 = [(0.40, 48.33, 21.6), (0.25, 86.3, 26.8), (0.17, 24.0, 6.0), (0.18, -1.0, -1.0)]
TYPES_CONFIG4 = [(0.01, 24.0, 4.0), (0.29, 86.3, 26.8), (0.70, -1.0, -1.0)]


@dataclass
class Workload:
    name: str
    n_cells: int
    types: np.ndarray            # [n_sets][n_types][3]
    t_max: float
    phi: float                   # 0.0 -> default (min non-empty bin)
    track_ratio: bool
    seed: int
    values: np.ndarray = field(repr=False, default=None)
    freqs: np.ndarray = field(repr=False, default=None)


def _min_nonempty(values, freqs) -> float:
    return float(values[freqs > 0].min())


def sweep_types(n_sets: int = 1024) -> np.ndarray:
    """Config 5: p1 in 8 x mu1 in 8 x sigma1 in 4 x mu2 in 4 values around config 1's fit; the quiescent
    type takes the remainder of the proportions."""
    p1s = np.linspace(0.40, 0.61, 8)
    mu1s = np.linspace(40.0, 57.5, 8)
    sd1s = np.linspace(15.0, 27.0, 4)
    mu2s = np.linspace(76.0, 97.0, 4)
    out = []
    for p1 in p1s:
        for mu1 in mu1s:
            for sd1 in sd1s:
                for mu2 in mu2s:
                    p2 = 0.29
                    out.append([(p1, mu1, sd1), (p2, mu2, 26.8), (1.0 - p1 - p2, -1.0, -1.0)])
    arr = np.array(out, dtype=np.float64)
    assert arr.shape[0] == 1024
    return arr[:n_sets]


def workload(config: int, scale: float = 1.0) -> Workload:
    """The five BASELINE.json configs; `scale` shrinks the cell count (parity tests use small scales)."""
    if config == 1:
        n = max(1, int(1e4 * scale))
        v, f = synthetic_histogram(n)
        return Workload("config1", n, np.array([TYPES_CONFIG1]), 168.0, _min_nonempty(v, f), False,
                        0x5EED0001, v, f)
    if config == 2:
        n = max(1, int(1e6 * scale))
        v, f = synthetic_histogram(n)
        return Workload("config2", n, np.array([TYPES_CONFIG2]), 240.0, 0.5, True, 0x5EED0002, v, f)
    if config == 3:
        n = max(1, int(1e8 * scale))
        v, f = synthetic_histogram(n)
        return Workload("config3", n, np.array([TYPES_CONFIG2]), 336.0, 0.0, False, 0x5EED0003, v, f)
    if config == 4:
        n = max(1, int(1e4 * scale))
        v, f = synthetic_histogram(n)
        return Workload("config4", n, np.array([TYPES_CONFIG4]), 720.0, 1e-7, False, 0x5EED0004, v, f)
    if config == 5:
        n = max(1, int(1e6 * scale))
        v, f = synthetic_histogram(n)
        return Workload("config5", n, sweep_types(1024), 168.0, 0.5, False, 0x5EED0005, v, f)
    raise ValueError("config must be 1..5")


def expected_depth(t_max: float, mean: float) -> float:
    return math.inf if mean <= 0 else t_max / mean
