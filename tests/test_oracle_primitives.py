"""Pin the explicit primitive restatements (what the CUDA kernels implement) against numpy / scipy."""
import numpy as np
import pytest
from scipy.ndimage import gaussian_filter

from oracle import primitives as P


@pytest.mark.parametrize("n", [0, 1, 3, 7, 8, 9, 50, 97, 128, 129, 150, 163, 179, 180, 257, 360, 4356, 6084, 21870])
def test_pairwise_sum_matches_numpy(n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 3, n)
    assert P.pairwise_sum(a) == np.sum(a)
    # leaves + tree give the same recursion
    if n > 0:
        leaves = P.pairwise_leaves(n)
        assert sum(l for _, l in leaves) == n and all(0 < l <= 128 for _, l in leaves)


def test_pairwise_sum_is_what_axis2_and_3d_reductions_use():
    rng = np.random.default_rng(1)
    vol = rng.standard_normal((36, 13, 13))
    assert P.pairwise_sum(vol.reshape(-1)) == np.sum(vol)
    g = rng.standard_normal((11, 11, 163))
    ref = np.sum(g, axis=2)
    mine = np.array([[P.pairwise_sum(g[a, b]) for b in range(11)] for a in range(11)])
    assert np.array_equal(ref, mine)


@pytest.mark.parametrize("sigma,shape", [(0.4, (101, 101)), (2, (251, 249)), (1.0, (121, 121)), (2, (40, 9)), (0.4, (3, 7))])
def test_gaussian_filter_bit_exact(sigma, shape):
    rng = np.random.default_rng(3)
    x = np.where(rng.random(shape) < 0.1, 0.0, np.log(0.15))
    assert np.array_equal(P.gaussian_filter_reflect(x, sigma), gaussian_filter(x, sigma=sigma))
    y = rng.standard_normal(shape)
    assert np.array_equal(P.gaussian_filter_reflect(y, sigma), gaussian_filter(y, sigma=sigma))


def test_unique_xy_matches_numpy_rows():
    rng = np.random.default_rng(5)
    xi, yi = rng.integers(0, 60, 180), rng.integers(0, 60, 180)
    ref = np.unique(np.column_stack((xi, yi)), axis=0)
    ux, uy = P.unique_xy(xi, yi)
    assert np.array_equal(ref[:, 0], ux) and np.array_equal(ref[:, 1], uy)


@pytest.mark.parametrize("n,size", [(1000, 1), (10, 10), (6084, 1), (64, 64)])
def test_legacy_choice(n, size):
    rng = np.random.default_rng(n)
    p = rng.random(n) ** 6
    p /= p.sum()
    np.random.seed(11)
    ref = np.random.choice(np.arange(n), size, p=p)
    np.random.seed(11)
    u = np.random.random_sample(size)
    assert np.array_equal(P.legacy_choice_index(p, u), ref)


def test_linspace():
    for th in (-0.463373, 1.25, 31.29101329218997):
        a = P.linspace(th - np.pi / 2, th + np.pi / 2, 180)
        assert np.array_equal(a, np.linspace(th - np.pi / 2, th + np.pi / 2, num=180))


def test_division_by_a_fixed_cell_size_with_a_hoisted_reciprocal_is_ieee_division():
    """csrc/match.cu `ddiv_rcp`: the index division (q - begin) / unitLength (ScanMatcher_OGBased.py:174-175) uses the
    correctly rounded reciprocal r = RN(1 / b) once per list: q = RN(a r), rem = a - b q (exact in one FMA),
    RN(q + rem r).  Markstein's theorem says this IS the correctly rounded quotient; checked here in exact rational
    arithmetic (fractions -> float conversion rounds to nearest even) against Python's IEEE division."""
    from fractions import Fraction
    rng = np.random.default_rng(11)

    def fma(x, y, z):
        return float(Fraction(x) * Fraction(y) + Fraction(z))

    divisors = [0.02, 0.05, 0.1, 0.25, 0.5, 0.05 * 5, 0.02 * 5, 0.1 * 2] + list(rng.uniform(0.01, 1.0, 8))
    for b in divisors:
        r = 1.0 / b
        cells = rng.integers(0, 1300, 700)
        samples = np.concatenate([rng.uniform(0.0, 70.0, 700),
                                  cells * b,                                   # exact multiples of the cell size
                                  cells * b + rng.choice([-1, 1], 700) * np.spacing(cells * b)])
        for a in samples:
            a = float(a)
            q = a * r
            rem = fma(-b, q, a)
            got = fma(rem, r, q)
            assert got == a / b, (a, b, got, a / b)
