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
                      ctr=(env & 0xffffffff, env >> 32, step, (domain << 24) | (slot >> 2))
                     )[slot & 3]

``env`` is the GLOBAL env index (shard-invariant), ``step`` the caller's step counter,
``domain`` 0 for step(), 1 for reset().
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

DOMAIN_STEP = 0
DOMAIN_RESET = 1


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


def draw_block(seed, env, step, domain, block):
    """One 4-word Philox block per env; env may be an array of global indices."""
    env = np.asarray(env, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return philox4x32_10(env & MASK, env >> np.uint64(32),
                         np.uint64(int(step) & 0xFFFFFFFF),
                         np.uint64(((int(domain) & 0xFF) << 24) | (int(block) & 0xFFFFFF)),
                         seed & 0xFFFFFFFF, seed >> 32)


def draw_slots(seed, env, step, domain, n_slots):
    """uint32 array [len(env), n_slots] of draw words, slot-major per the contract."""
    env = np.atleast_1d(np.asarray(env, dtype=np.uint64))
    out = np.empty((env.shape[0], n_slots), dtype=np.uint32)
    for b in range((n_slots + 3) // 4):
        words = draw_block(seed, env, step, domain, b)
        for j in range(4):
            s = 4 * b + j
            if s < n_slots:
                out[:, s] = words[j]
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
