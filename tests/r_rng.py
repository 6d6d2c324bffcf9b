"""R's default random number stream (Mersenne-Twister + inversion normals),
restated so the RNG-dependent golden values of the reference's saved test
output (tests/all.Rout.save, set.seed(42); tests/Examples/*.Rout.save,
set.seed(1) via cleanEx()) can be regenerated without R.

Test infrastructure only.
"""
import numpy as np
from scipy.special import ndtri


class RRng:
    def __init__(self, seed: int):
        s = np.uint32(seed)
        with np.errstate(over="ignore"):
            for _ in range(50):  # initial scrambling
                s = np.uint32(69069) * s + np.uint32(1)
            key = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = np.uint32(69069) * s + np.uint32(1)
                key[j] = s
        # FixupSeeds: dummy[0] = mti = 624, the 624 words after it are the state
        bg = np.random.MT19937()
        st = bg.state
        st["state"]["key"] = key[1:].copy()
        st["state"]["pos"] = 624
        bg.state = st
        self.bg = bg

    def unif_rand(self, n=None):
        raw = self.bg.random_raw(1 if n is None else n).astype(np.float64)
        u = raw * 2.3283064365386963e-10
        lo = 2.328306437080797e-10
        u = np.where(u <= 0.0, 0.5 * lo, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * lo, u)
        return float(u[0]) if n is None else u

    def runif(self, n, a=0.0, b=1.0):
        u = self.unif_rand(n)
        return a + (b - a) * u

    def rnorm(self, n):
        BIG = 134217728.0
        u = self.unif_rand(2 * n)
        v = np.floor(BIG * u[0::2]) + u[1::2]
        return ndtri(v / BIG)
