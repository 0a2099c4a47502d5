"""Test infrastructure: pure-Python restatement of what `np.random.default_rng(seed)` does for the draws CrowdSimPlus.reset makes
(crowd_sim_plus/envs/crowd_sim_plus.py:658-664, 454-481, 522-605 use only `rng.random()` and `rng.uniform(a, b)`):
numpy's SeedSequence entropy hashing, PCG64 (XSL-RR 128/64) seeding and stream, `random() = (u64 >> 11) * 2**-53`,
`uniform(a, b) = a + (b - a) * random()`.  csrc/scene_kernels.cu implements the same arithmetic on the device; this file is the
CPU statement of it, pinned against numpy itself in tests/test_pcg64_oracle_cpu.py.  Only tests may import it."""

M32 = 0xFFFFFFFF
M64 = (1 << 64) - 1
M128 = (1 << 128) - 1
INIT_A, MULT_A, INIT_B, MULT_B = 0x43B0D7E5, 0x931E8875, 0x8B51F9DD, 0x58F38DED
MIX_MULT_L, MIX_MULT_R, XSHIFT = 0xCA01F9DD, 0x4973F715, 16
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645


def seed_sequence_state(seed, n_words32=8):
    """SeedSequence(seed).generate_state(n_words32, uint32) for a non-negative int seed."""
    ent = []
    x = int(seed)
    if x == 0:
        ent = [0]
    while x > 0:
        ent.append(x & M32)
        x >>= 32
    hc = [INIT_A]

    def hashmix(v):
        v = (v ^ hc[0]) & M32
        hc[0] = (hc[0] * MULT_A) & M32
        v = (v * hc[0]) & M32
        return v ^ (v >> XSHIFT)

    def mix(a, b):
        r = (MIX_MULT_L * a - MIX_MULT_R * b) & M32
        return r ^ (r >> XSHIFT)

    pool = [hashmix(ent[i] if i < len(ent) else 0) for i in range(4)]
    for i_src in range(4):
        for i_dst in range(4):
            if i_src != i_dst:
                pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src]))
    for i_src in range(4, len(ent)):
        for i_dst in range(4):
            pool[i_dst] = mix(pool[i_dst], hashmix(ent[i_src]))
    hb = INIT_B
    out = []
    for i in range(n_words32):
        v = pool[i % 4] ^ hb
        hb = (hb * MULT_B) & M32
        v = (v * hb) & M32
        out.append(v ^ (v >> XSHIFT))
    return out


class Pcg64:
    def __init__(self, seed):
        w = seed_sequence_state(seed, 8)
        u = [w[2 * i] | (w[2 * i + 1] << 32) for i in range(4)]
        initstate, initseq = (u[0] << 64) | u[1], (u[2] << 64) | u[3]
        self.inc = ((initseq << 1) | 1) & M128
        self.state = 0
        self._step()
        self.state = (self.state + initstate) & M128
        self._step()

    def _step(self):
        self.state = (self.state * PCG_MULT + self.inc) & M128

    def next64(self):
        self._step()
        hi, lo = self.state >> 64, self.state & M64
        x, rot = hi ^ lo, self.state >> 122
        return ((x >> rot) | (x << ((64 - rot) & 63))) & M64

    def random(self):
        return (self.next64() >> 11) * (1.0 / 9007199254740992.0)

    def uniform(self, a, b):
        return a + (b - a) * self.random()
