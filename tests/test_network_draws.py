"""Network's joint failure draw (include/pomdp_b200.h, "Network draws"; network.py:94-99): the alias table every
implementation of the contract builds, the exact distribution it samples, and the per-machine words the oracle's
one-binomial-per-machine step consumes."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import philox

PAIRS = [(.1, .33), (.33, .1), (0., 1.), (1., 1.), (0., 0.), (.5, .5), (1., 0.), (.25, 0.), (1e-9, 1 - 1e-9), (.999, .9991)]


@pytest.mark.parametrize("p,q", PAIRS)
def test_python_and_c_build_the_same_table(p, q):
    thr, al = C.network_alias(p, q)
    thr2, al2 = philox.network_alias(philox.bern_T(p), philox.bern_T(q))
    assert thr.tolist() == thr2 and al.tolist() == al2
    assert all(0 <= t < (1 << 24) for t in thr2)
    assert all(a < philox.NET_CODES for k, a in enumerate(al2) if a != k)      # only real outcomes are alias targets


@pytest.mark.parametrize("p,q", PAIRS)
def test_alias_table_samples_the_product_of_the_reference_bernoullis(p, q):
    """In exact rational arithmetic: the probability of every joint outcome under the table vs. the product of the
    per-machine probabilities [u < lo], [lo <= u < hi], [u >= hi] (u uniform on 2^32 words)."""
    T_p, T_q = philox.bern_T(p), philox.bern_T(q)
    lo, hi = min(T_p, T_q), max(T_p, T_q)
    c = [Fraction(lo, 1 << 32), Fraction(hi - lo, 1 << 32), Fraction((1 << 32) - hi, 1 << 32)]
    dist = philox.network_alias_distribution(T_p, T_q)
    assert sum(dist) == 1
    worst = Fraction(0)
    for k in range(philox.NET_COLS):
        exact, r = Fraction(int(k < philox.NET_CODES)), k
        for _ in range(philox.NET_GROUP):
            exact *= c[r % 3]
            r //= 3
        if exact == 0:
            assert dist[k] == 0, k                      # impossible outcomes stay impossible (p or q exactly 0 or 1)
        worst = max(worst, abs(dist[k] - exact))
    # every column's 24-bit threshold is off by < 2^-24 of the column's 1/256: < 2^-32 per column, 256 columns at most
    assert worst < Fraction(256, 1 << 32), float(worst)
    if (p, q) == (.1, .33):
        assert worst < Fraction(1, 10 ** 8), float(worst)          # 7.5e-9 at the reference's own probabilities
    # the marginals the reference's binomial(1, p) / binomial(1, q) have, for every machine of the group
    for i in range(philox.NET_GROUP):
        m_lo = sum(dist[k] for k in range(philox.NET_CODES) if (k // 3 ** i) % 3 == 0)
        m_hi = sum(dist[k] for k in range(philox.NET_CODES) if (k // 3 ** i) % 3 <= 1)
        assert abs(m_lo - c[0]) < Fraction(256, 1 << 32) and abs(m_hi - c[0] - c[1]) < Fraction(256, 1 << 32)


@pytest.mark.parametrize("n", [4, 7, 10, 19, 30])
def test_per_machine_words_python_equals_c_and_decide_like_the_digits(n):
    env = np.arange(4093, 4093 + 3000)
    for p, q in [(.1, .33), (.33, .1), (1., 0.), (0., 0.)]:
        d = philox.network_draws(77, env, 5, n, p, q)
        assert np.array_equal(d, C.network_draws(77, 4093, 3000, 5, n, p, q))
        T_p, T_q = philox.bern_T(p), philox.bern_T(q)
        G = (n + 4) // 5
        w = philox.draw_slots(77, env, 5, philox.DOMAIN_STEP, G + 1)
        digits = philox.network_digits(w[:, :G], T_p, T_q)[:, :n]
        u = d[:, :n].astype(np.float64) / 2.0 ** 32                       # the oracle compares word / 2^32 < p
        assert np.array_equal(u < min(p, q), digits == 0) and np.array_equal(u < max(p, q), digits <= 1)
        assert np.array_equal(d[:, n], w[:, G])                            # the observation draw's own word


def test_failure_frequencies_at_the_reference_probabilities():
    d = philox.network_draws(3, np.arange(400000), 9, 10)
    u = d[:, :10].astype(np.float64) / 2.0 ** 32
    for thr in (.1, .33):
        f = (u < thr).mean(0)
        assert np.abs(f - thr).max() < 5 * np.sqrt(thr * (1 - thr) / 400000), f
    # machines of one env are independent: pairwise covariance of the failure indicators ~ 0
    x = (u < .33).astype(np.float64)
    cov = np.cov(x.T)
    off = cov - np.diag(np.diag(cov))
    assert np.abs(off).max() < 5 * .33 * .67 / np.sqrt(400000)
