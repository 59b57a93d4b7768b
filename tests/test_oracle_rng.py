"""The oracle's restatement of numpy's legacy RandomState (MT19937, polar gauss, masked randint) vs numpy itself."""
import numpy as np
import pytest

from oracle.oracle import NumpyLegacyRng


@pytest.mark.parametrize("seed", [0, 1, 12345, 123456789, 2 ** 32 - 1])
def test_mixed_draws_bit_equal(seed):
    r, n = NumpyLegacyRng(seed), np.random.RandomState(seed)
    for i in range(4000):
        k = i % 7
        if k == 0:
            a, b = r.standard_normal(), n.standard_normal()
        elif k == 1:
            a, b = r.standard_exponential(), n.standard_exponential()
        elif k == 2:
            a, b = r.random_sample(), n.random_sample()
        elif k == 3:
            a, b = r.randint(0, 100), n.randint(0, 100)
        elif k == 4:
            a, b = r.randint(0, 2 ** 32), int(n.randint(0, 2 ** 32, dtype="uint64"))
        elif k == 5:
            a, b = r.randint(0, 2), n.randint(0, 2)
        else:
            a, b = r.randint(250, 501), n.randint(250, 501)
        assert a == b, (i, k, a, b)


def test_scaled_identities():
    # normal(loc, scale) == loc + scale * standard_normal(), etc.
    a, b = np.random.RandomState(7), np.random.RandomState(7)
    for _ in range(1000):
        assert a.normal(3.5, 1000.0) == 3.5 + 1000.0 * b.standard_normal()
        assert a.exponential(1e12) == b.standard_exponential() * 1e12
        assert a.uniform(0.05, 1.0) == 0.05 + (1.0 - 0.05) * b.random_sample()
        assert a.choice(6, 1, [.25, .25, .2, .15, .1, .05])[0] == b.randint(0, 6)
