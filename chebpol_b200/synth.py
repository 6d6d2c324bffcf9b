"""Counter-based synthetic points: coordinate d of global point i is a pure
function of (seed, i*rank + d).  ``points`` is the numpy twin of the device generator
``chebpol_cuda_synth_points`` (csrc/synth.cu), bit for bit, so a CPU slice and a GPU slice of the
same batch hold identical values however the batch is split over GPUs."""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def points(first_point: int, npts: int, rank: int, seed: int = 42, lo: float = -1.0, hi: float = 1.0) -> np.ndarray:
    """(npts, rank) C-contiguous float64: global points first_point .. first_point + npts - 1."""
    with np.errstate(over="ignore"):
        e = np.arange(npts * rank, dtype=np.uint64) + np.uint64((first_point * rank) & 0xFFFFFFFFFFFFFFFF)
        z = _splitmix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + e * _GOLD)
    u = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return np.ascontiguousarray((lo + (hi - lo) * u).reshape(npts, rank))


def points_at(index, rank: int, seed: int = 42, lo: float = -1.0, hi: float = 1.0) -> np.ndarray:
    """Rows of the batch at arbitrary global point indices (strided parity samples)."""
    idx = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        e = idx[:, None] * np.uint64(rank) + np.arange(rank, dtype=np.uint64)[None, :]
        z = _splitmix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + e * _GOLD)
    u = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return np.ascontiguousarray(lo + (hi - lo) * u)
