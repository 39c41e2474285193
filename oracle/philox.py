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
