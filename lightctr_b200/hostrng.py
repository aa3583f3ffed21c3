"""Host-side RNG of the reference, kept on the host on purpose (init parity is free if the
glibc rand() stream is consumed in the reference's own call order).

glibc srand/rand (random_r TYPE_3 additive-feedback generator) + util/random.h's UniformNumRand (:21),
UniformNumRand2 (:25), GaussRand (:42-58, polar Box-Muller with a cached second deviate), SampleBinary (:82).
"""
import math

import numpy as np

RAND_MAX = 2147483647


class GlibcRand:
    def __init__(self, seed=1):
        self.srand(seed)
        self._phase = 0
        self._v1 = self._v2 = self._s = 0.0

    def srand(self, seed):
        seed = int(seed) & 0xFFFFFFFF
        if seed == 0:
            seed = 1
        r = [0] * 344
        r[0] = seed if seed < 2 ** 31 else seed - 2 ** 32
        for i in range(1, 31):
            hi, lo = divmod(abs(r[i - 1]), 127773) if r[i - 1] >= 0 else (-(abs(r[i - 1]) // 127773), -(abs(r[i - 1]) % 127773))
            w = 16807 * lo - 2836 * hi
            if w < 0:
                w += 2147483647
            r[i] = w
        for i in range(31, 34):
            r[i] = r[i - 31]
        for i in range(34, 344):
            r[i] = (r[i - 31] + r[i - 3]) & 0xFFFFFFFF
        self._ring = [x & 0xFFFFFFFF for x in r[310:344]]
        self._i = 0

    def rand(self):
        i = self._i
        v = (self._ring[(i + 3) % 34] + self._ring[(i + 31) % 34]) & 0xFFFFFFFF
        self._ring[i] = v
        self._i = (i + 1) % 34
        return v >> 1

    def rand_array(self, n):
        """n consecutive rand() values (vectorised in blocks of 3, the generator's shortest lag)."""
        out = np.empty(n, np.int64)
        ring = self._ring
        i = self._i
        for j in range(n):
            v = (ring[(i + 3) % 34] + ring[(i + 31) % 34]) & 0xFFFFFFFF
            ring[i] = v
            i = (i + 1) % 34
            out[j] = v >> 1
        self._i = i
        return out

    def uniform(self):  # UniformNumRand [0,1)
        return self.rand() / (RAND_MAX + 1.0)

    def uniform2(self):  # UniformNumRand2 (0,1)
        return (self.rand() + 1.0) / (RAND_MAX + 2.0)

    def sample_binary(self, p):
        return self.uniform() < p

    def gauss(self):
        if self._phase == 0:
            while True:
                self._v1 = 2.0 * self.uniform2() - 1.0
                self._v2 = 2.0 * self.uniform2() - 1.0
                self._s = self._v1 * self._v1 + self._v2 * self._v2
                if not (self._s >= 1.0 or self._s == 0.0):
                    break
            x = self._v1 * math.sqrt(-2.0 * math.log(self._s) / self._s)
        else:
            x = self._v2 * math.sqrt(-2.0 * math.log(self._s) / self._s)
        self._phase = 1 - self._phase
        return x

    def gauss_fill(self, n, factor_cnt):
        """V[i] = GaussRand() * (float)(1.0/sqrt(k))  (fm_algo_abst.h:62-65), vectorised polar Box-Muller."""
        scale = float(np.float32(1.0 / math.sqrt(factor_cnt)))
        out = np.empty(n, np.float64)
        filled = 0
        if self._phase == 1 and n > 0:
            out[0] = self._v2 * math.sqrt(-2.0 * math.log(self._s) / self._s)
            self._phase = 0
            filled = 1
        while filled < n:
            need_pairs = (n - filled + 1) // 2
            m = int(need_pairs * 1.35) + 16
            # draw 2*m uniforms; accept/reject sequentially preserves the stream only if we consume exactly
            # what the scalar code would: do it pairwise and stop as soon as enough pairs are accepted
            raw = self.rand_array(2 * m)
            u = (raw + 1.0) / (RAND_MAX + 2.0)
            v1 = 2.0 * u[0::2] - 1.0
            v2 = 2.0 * u[1::2] - 1.0
            s = v1 * v1 + v2 * v2
            ok = ~((s >= 1.0) | (s == 0.0))
            idx = np.nonzero(ok)[0]
            if len(idx) >= need_pairs:
                last = idx[need_pairs - 1]
                # un-consume the draws after pair `last`
                unused = 2 * m - 2 * (last + 1)
                self._rewind(raw, unused)
                idx = idx[:need_pairs]
            f = np.sqrt(-2.0 * np.log(s[idx]) / s[idx])
            x1, x2 = v1[idx] * f, v2[idx] * f
            pair = np.empty(2 * len(idx), np.float64)
            pair[0::2], pair[1::2] = x1, x2
            take = min(len(pair), n - filled)
            out[filled:filled + take] = pair[:take]
            filled += take
            if take < len(pair):  # odd count: second deviate of the last pair stays cached
                self._phase = 1
                self._v1, self._v2, self._s = float(v1[idx[-1]]), float(v2[idx[-1]]), float(s[idx[-1]])
        return (out * scale).astype(np.float32)

    def _rewind(self, raw, unused):
        # stepping the additive-feedback generator backwards: ring[i] was overwritten with the new value;
        # the overwritten (oldest) value is recoverable as v[-34] = v[-3] - v[-31]... simpler: restore by replay.
        if unused == 0:
            return
        # replay: we know the generator is deterministic; reconstruct the state by reversing `unused` steps.
        ring, i = self._ring, self._i
        for _ in range(unused):
            i = (i - 1) % 34
            # new = old[i+3] + old[i+31]; after the step ring[i] = new.  The value that was in ring[i] before
            # (call it o) satisfies: the element 31 steps later: ring[(i+31)%34]_new = ring[(i+31+3)%34] + o ... too
            # indirect -- instead use: v[n] = v[n-31] + v[n-3]  =>  v[n-31] = v[n] - v[n-3].
            # ring[i] currently holds v[n]; ring[(i-3)%34] holds v[n-3]; the slot must be restored to v[n-34].
            # v[n-34] = v[n-3] - v[n-31-... ] is not directly available, but v[n-34+31] = v[n-3] = v[n-34] + v[n-6]
            #   => v[n-34] = v[n-3] - v[n-6].
            ring[i] = (ring[(i - 3) % 34] - ring[(i - 6) % 34]) & 0xFFFFFFFF
        self._i = i
