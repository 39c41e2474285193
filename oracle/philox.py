"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11), numpy restatement.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): imported by tests/, by
``__graft_entry__.smoke()`` and by ``bench.py``'s cpu_baseline / ``--impl reference``
leg.  The product (gym_pomdp_b200/) never imports anything under oracle/.

The reference (d3sm0/gym_pomdp) draws from numpy's global MT19937
(``np.random.binomial/uniform/randint/choice``; e.g. rock.py:80, rock.py:404,
tag.py:204-205, network.py:94-112, battleship.py:36, coord.py:68).  Its bit stream is
not part of the parity contract (BASELINE.json north_star: distributional parity);
the CUDA kernels use Philox instead, and the *coupled* parity tests feed exactly
these Philox words into the unmodified reference through ``oracle/ref_shim.py``.

Draw-slot contract shared by the kernels, the C oracle and this file::

    word(seed, env, step, domain, slot) =
        philox4x32_10(key=(seed & 0xffffffff, seed >> 32),
                      ctr=(g & 0xffffffff, g >> 32, step, (domain << 24) | slot)
                     )[env & 3]          with g = env >> 2

i.e. one Philox block holds the SAME slot of FOUR consecutive envs (a GPU thread owns an
aligned group of four envs and pays one Philox call per slot).  ``env`` is the GLOBAL env
index (shard-invariant), ``step`` the caller's step counter, ``domain`` 0 for step(), 1
for reset().
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

DOMAIN_STEP = 0
DOMAIN_RESET = 1
DOMAIN_POLICY = 2


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays (broadcast); returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK
    c1 = np.asarray(c1, dtype=np.uint64) & MASK
    c2 = np.asarray(c2, dtype=np.uint64) & MASK
    c3 = np.asarray(c3, dtype=np.uint64) & MASK
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for _ in range(10):
        p0 = M0 * c0  # < 2^64, no overflow
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def draw_quad(seed, group, step, domain, slot):
    """One 4-word Philox block per draw group (= 4 consecutive envs); group may be an array."""
    group = np.asarray(group, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return philox4x32_10(group & MASK, group >> np.uint64(32),
                         np.uint64(int(step) & 0xFFFFFFFF),
                         np.uint64(((int(domain) & 0xFF) << 24) | (int(slot) & 0xFFFFFF)),
                         seed & 0xFFFFFFFF, seed >> 32)


def draw_slots(seed, env, step, domain, n_slots):
    """uint32 array [len(env), n_slots] of draw words per the contract above."""
    env = np.atleast_1d(np.asarray(env, dtype=np.uint64))
    out = np.empty((env.shape[0], n_slots), dtype=np.uint32)
    lane = (env & np.uint64(3)).astype(np.int64)
    rows = np.arange(env.shape[0])
    for s in range(n_slots):
        words = np.stack(draw_quad(seed, env >> np.uint64(2), step, domain, s), axis=1)
        out[:, s] = words[rows, lane]
    return out


DOMAIN_SHIP = 3


def draw_env_slots(seed, env, step, domain, n_slots):
    """uint32 [len(env), n_slots] for draws keyed by the ENV itself (BattleShip's fixed-time placement, domain SHIP):
    word(env, slot) = philox(key=seed, ctr=(lo32(env), hi32(env), step, domain << 24 | slot >> 2))[slot & 3] -- one block
    holds four consecutive SLOTS of one env."""
    env = np.atleast_1d(np.asarray(env, dtype=np.uint64))
    out = np.empty((env.shape[0], n_slots), dtype=np.uint32)
    for s in range(n_slots):
        out[:, s] = draw_quad(seed, env, step, domain, s >> 2)[s & 3]
    return out


# ---------------------------------------------------------------------------------------
# Network: the joint failure draw (include/pomdp_b200.h, "Network draws")
# ---------------------------------------------------------------------------------------
# The reference draws binomial(1, p or q) once per machine (network.py:94-99).  One machine is
# one uniform u against the two thresholds T_p, T_q -- three outcomes ("digits"): 0 = u < lo
# (fails either way), 1 = lo <= u < hi (fails only under the larger probability), 2 = stays up,
# with lo = min(T_p, T_q), hi = max.  The kernels sample the 3^5 joint outcomes of FIVE machines
# from ONE 32-bit word through a 256-column alias table built in integer arithmetic; slot g of
# the step domain decides machines 5g..5g+4, slot ceil(n/5) is the observation draw.
NET_GROUP = 5
NET_CODES = 243
NET_COLS = 256


def bern_T(p):
    """ceil(p * 2^32) clipped to [0, 2^32] (p * 2^32 is exact in a double)."""
    import math
    return max(0, min(1 << 32, math.ceil(p * 4294967296.0)))


def network_alias(T_p, T_q):
    """The alias columns: (thr24[256], alias[256]) with column k's own outcome = code k.

    Integer arithmetic only, so every implementation of the contract builds the same table:
    digit counts c = (lo, hi - lo, 2^32 - hi); weight of code k = sum d_i 3^i is the chained
    product ((c[d0] * c[d1] >> 32) * c[d2] >> 32) ...; Vose's pairing over V[k] = 256 W[k]
    against the column mean S = sum W, small (V < S) and large columns listed in ascending k
    and paired from the END of their lists; thr24 = floor(V * 2^24 / S)."""
    lo, hi = min(T_p, T_q), max(T_p, T_q)
    c = (lo, hi - lo, (1 << 32) - hi)
    V = [0] * NET_COLS
    for k in range(NET_CODES):
        a, r = 1 << 32, k
        for _ in range(NET_GROUP):
            a = (a * c[r % 3]) >> 32
            r //= 3
        V[k] = a
    S = sum(V)
    V = [v * NET_COLS for v in V]
    alias = list(range(NET_COLS))
    thr24 = [0] * NET_COLS
    small = [k for k in range(NET_COLS) if V[k] < S]
    large = [k for k in range(NET_COLS) if V[k] >= S]
    while small and large:
        s_, l_ = small.pop(), large.pop()
        thr24[s_] = (V[s_] << 24) // S
        alias[s_] = l_
        V[l_] -= S - V[s_]
        (small if V[l_] < S else large).append(l_)
    return thr24, alias


def network_alias_distribution(T_p, T_q):
    """Exact probability (a Fraction) of every code under the alias table: the distribution the kernels sample."""
    from fractions import Fraction
    thr24, alias = network_alias(T_p, T_q)
    q = [Fraction(0)] * NET_COLS
    for k in range(NET_COLS):
        f = Fraction(thr24[k], 1 << 24) if alias[k] != k else Fraction(1)
        q[k] += f / NET_COLS
        if alias[k] != k:
            q[alias[k]] += (1 - f) / NET_COLS
    return q


def network_digits(words, T_p, T_q):
    """uint32 [N, G] joint draw words -> int [N, 5 G] digits (machine 5g + i <- digit i of word g's code)."""
    thr24, alias = network_alias(T_p, T_q)
    thr = np.asarray(thr24, dtype=np.uint64) << np.uint64(8)
    own = np.arange(NET_COLS)
    al = np.asarray(alias)
    words = np.asarray(words, dtype=np.uint64)
    col = (words & np.uint64(NET_COLS - 1)).astype(np.int64)
    code = np.where((words < thr[col]) | (al[col] == own[col]), own[col], al[col])
    digits = np.stack([(code // 3 ** i) % 3 for i in range(NET_GROUP)], axis=-1)      # [N, G, 5]
    return digits.reshape(words.shape[0], -1)


def network_draws(seed, env, step, n_machines, p=0.1, q=0.33):
    """uint32 [len(env), n_machines + 1]: per-machine words that make ``word / 2^32 < p`` (resp. ``< q``) reproduce the
    joint draw's digits -- 0 for digit 0, lo for digit 1, hi for digit 2 -- followed by the observation draw's word.
    This is what the oracle's network_step (one binomial per machine, like the reference) consumes."""
    T_p, T_q = bern_T(p), bern_T(q)
    lo, hi = min(T_p, T_q), max(T_p, T_q)
    G = (n_machines + NET_GROUP - 1) // NET_GROUP
    w = draw_slots(seed, env, step, DOMAIN_STEP, G + 1)
    digits = network_digits(w[:, :G], T_p, T_q)[:, :n_machines]
    rep = np.array([0, lo, min(hi, 0xFFFFFFFF)], dtype=np.uint64)     # digit 2 has probability 0 when hi = 2^32
    out = np.empty((w.shape[0], n_machines + 1), dtype=np.uint32)
    out[:, :n_machines] = rep[digits].astype(np.uint32)
    out[:, n_machines] = w[:, G]
    return out


def kat():
    """Known-answer vectors of Random123 (kat_vectors, philox4x32 10 rounds)."""
    vecs = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, exp in vecs:
        got = tuple(int(np.atleast_1d(w)[0]) for w in philox4x32_10(*ctr, *key))
        assert got == exp, (ctr, key, [hex(g) for g in got])
    return True


if __name__ == "__main__":
    print("philox KAT", kat())
