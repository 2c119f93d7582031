"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Pure-Python restatement of the random-number layer the reference's hot path consumes:

* dSFMT-19937 (Saito & Matsumoto, "dSFMT" v2.1; vendored by the reference under src/dsfmt/,
  parameters in src/dsfmt/dSFMT-params19937.h, recursion in src/dsfmt/dSFMT.c:150-165,
  initialisation dSFMT.c:625-646, period certification dSFMT.c:440-470,
  `dsfmt_genrand_open_open` in src/dsfmt/dSFMT.h:341-355).
* RngWrapper (src/rngwrapper.h:43-119, src/rngwrapper.cpp:30-50): seed scrambling in uint32
  wrap-around arithmetic, rand01 / randRange / randInt.

Parity of this restatement is pinned in tests/test_oracle_vs_reference.py against the reference's
own dSFMT compiled into oracle/_ref (when present) and against tests/golden/rng_*.npz.
"""

import struct

_MASK64 = (1 << 64) - 1
_N = 191                      # (19937 - 128) / 104 + 1
_N64 = 2 * _N
_POS1 = 117
_SL1 = 19
_SR = 12
_MSK1 = 0x000FFAFFFFFFFB3F
_MSK2 = 0x000FFDFFFC90FFFD
_FIX1 = 0x90014964B32F4329
_FIX2 = 0x3B8D12AC548A7C7A
_PCV1 = 0x3D84E1AC0DC82880
_PCV2 = 0x0000000000000001
_LOW_MASK = 0x000FFFFFFFFFFFFF
_HIGH_CONST = 0x3FF0000000000000


class Dsfmt19937:
    """State: 191 128-bit words + 1 'lung' word, kept as a flat list of 384 uint64."""

    def __init__(self, seed32):
        seed32 &= 0xFFFFFFFF
        n32 = (_N + 1) * 4
        w = [0] * n32
        w[0] = seed32
        for i in range(1, n32):
            prev = w[i - 1]
            w[i] = (1812433253 * (prev ^ (prev >> 30)) + i) & 0xFFFFFFFF
        # little-endian pairing of 32-bit words into 64-bit words
        self.u = [w[2 * i] | (w[2 * i + 1] << 32) for i in range(n32 // 2)]
        for i in range(_N64):
            self.u[i] = (self.u[i] & _LOW_MASK) | _HIGH_CONST
        self._period_certification()
        self.idx = _N64

    def _period_certification(self):
        t0 = self.u[_N64] ^ _FIX1
        t1 = self.u[_N64 + 1] ^ _FIX2
        inner = (t0 & _PCV1) ^ (t1 & _PCV2)
        i = 32
        while i > 0:
            inner ^= inner >> i
            i >>= 1
        if inner & 1:
            return
        # PCV2 has its lowest bit set: flip that bit of the lung's second word
        self.u[_N64 + 1] ^= 1

    def _gen_rand_all(self):
        u = self.u
        l0, l1 = u[_N64], u[_N64 + 1]
        for i in range(_N):
            j = i + _POS1
            if j >= _N:
                j -= _N
            a0, a1 = u[2 * i], u[2 * i + 1]
            b0, b1 = u[2 * j], u[2 * j + 1]
            n0 = ((a0 << _SL1) & _MASK64) ^ (l1 >> 32) ^ ((l1 << 32) & _MASK64) ^ b0
            n1 = ((a1 << _SL1) & _MASK64) ^ (l0 >> 32) ^ ((l0 << 32) & _MASK64) ^ b1
            l0, l1 = n0, n1
            u[2 * i] = (l0 >> _SR) ^ (l0 & _MSK1) ^ a0
            u[2 * i + 1] = (l1 >> _SR) ^ (l1 & _MSK2) ^ a1
        u[_N64], u[_N64 + 1] = l0, l1

    def genrand_open_open(self):
        if self.idx >= _N64:
            self._gen_rand_all()
            self.idx = 0
        r = self.u[self.idx] | 1
        self.idx += 1
        return struct.unpack("<d", struct.pack("<Q", r))[0] - 1.0


def scramble_seed(seed, process_index):
    """rngwrapper.cpp:43 -- all arithmetic in uint32 wrap-around."""
    m32 = 0xFFFFFFFF
    a = (seed * 181) & m32
    b = (((process_index - 83) & m32) * 359) & m32
    return ((a * b) & m32) % 104729


class RngOracle:
    """Restatement of RngWrapper (rngwrapper.h:43-119)."""

    def __init__(self, seed=0, process_index=0):
        self.seed = seed
        self.process_index = process_index
        self.my_seed = scramble_seed(seed, process_index)
        self.gen = Dsfmt19937(self.my_seed)
        self.draws = 0

    def rand01(self):
        self.draws += 1
        return self.gen.genrand_open_open()

    def rand_range(self, low, high):
        return low + (high - low) * self.rand01()

    def rand_int(self, low, high):
        return low + int((high - low + 1.0) * self.rand01())
