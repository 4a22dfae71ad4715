"""oracle/orc_math.h (explicit-IEEE-sequence elementary functions) against libm / numpy: they replace
log / atan2 / sin / cos / pow in the oracle so that the GPU can match it bit for bit, and must themselves
stay within a few ulp of the correctly rounded values.  CPU only."""
import numpy as np

from oracle import oracle as O


def ulps(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_two_atanh_and_log():
    rng = np.random.default_rng(0)
    z = np.concatenate([rng.uniform(0, 0.2, 200000), rng.uniform(0.2, 0.999, 200000), 10.0 ** rng.uniform(-8, -1, 50000)])
    assert ulps(O.math_probe("two_atanh", z), 2 * np.arctanh(z)).max() <= 6
    x = np.concatenate([rng.uniform(1.0, 4.0, 200000), 10.0 ** rng.uniform(-300, 300, 100000)])
    got, ref = O.math_probe("log", x), np.log(x)
    far = np.abs(ref) > 0.05
    assert ulps(got[far], ref[far]).max() <= 3
    assert np.abs(got - ref).max() / np.maximum(np.abs(ref), 0.05).min() < 1  # finite everywhere
    assert np.abs(got[~far] - ref[~far]).max() < 4e-17


def test_atan2_all_quadrants_and_small_series():
    rng = np.random.default_rng(1)
    y, x = rng.normal(size=400000) * 10.0 ** rng.uniform(-3, 6, 400000), rng.normal(size=400000) * 10.0 ** rng.uniform(-3, 6, 400000)
    got, ref = O.math_probe("atan2", y, x), np.arctan2(y, x)
    assert ulps(got, ref).max() <= 4
    # axes and the branch cut
    yy = np.array([0.0, 1.0, 0.0, -1.0, 1.0, -1.0, 0.0])
    xx = np.array([1.0, 0.0, -1.0, 0.0, 1.0, -1.0, 0.0])
    assert np.allclose(O.math_probe("atan2", yy, xx), np.arctan2(yy, xx), rtol=0, atol=3e-16)
    q = rng.uniform(-0.1, 0.1, 200000)
    assert ulps(O.math_probe("atan_series8", q), np.arctan(q)).max() <= 2


def test_minimax_polynomials_of_the_common_path():
    """2 atanh(z) on z^2 <= 0.04 (7 coefficients) and 2 atan(t) on t^2 <= 0.01 (5): within 2 ulp of libm"""
    rng = np.random.default_rng(7)
    z = np.concatenate([rng.uniform(-0.2, 0.2, 300000), 10.0 ** rng.uniform(-9, -1, 50000), [0.2, -0.2, 0.0]])
    got, ref = O.math_probe("two_atanh_mm", z), 2 * np.arctanh(z)
    nz = ref != 0
    assert ulps(got[nz], ref[nz]).max() <= 2 and np.all(got[~nz] == 0)
    t = np.concatenate([rng.uniform(-0.1, 0.1, 300000), 10.0 ** rng.uniform(-9, -1, 50000), [0.1, -0.1, 0.0]])
    got, ref = O.math_probe("two_atan_mm", t), 2 * np.arctan(t)
    nz = ref != 0
    assert ulps(got[nz], ref[nz]).max() <= 2 and np.all(got[~nz] == 0)


def test_tier_polynomials_of_the_row_kernels():
    """far tier (4 coefficients, z^2, t^2 <= 0.0025), middle tier (5, z^2 <= 0.01) and the wide 2 atanh of the remaining
    rows (11, z^2 <= 0.16): within 2 ulp of libm on the range each tier is used on (the range tests sit just inside)"""
    rng = np.random.default_rng(8)
    small = 10.0 ** rng.uniform(-9, -2, 50000)
    for fn, lim, ref_fn in (("two_atanh_far", 0.05, np.arctanh), ("two_atan_far", 0.05, np.arctan),
                            ("two_atanh_mid", 0.1, np.arctanh), ("two_atanh_wide", 0.4, np.arctanh)):
        x = np.concatenate([rng.uniform(-lim, lim, 300000), small, [lim, -lim, 0.0]])
        got, ref = O.math_probe(fn, x), 2 * ref_fn(x)
        nz = ref != 0
        assert ulps(got[nz], ref[nz]).max() <= 2, fn
        assert np.all(got[~nz] == 0), fn
    # against 40-digit arithmetic at the upper end of the wide range, where the 11 coefficients matter most
    import mpmath
    mpmath.mp.dps = 40
    z = np.concatenate([rng.uniform(0.3, 0.4, 300), [0.4]])
    exact = np.array([float(2 * mpmath.atanh(mpmath.mpf(float(v)))) for v in z])
    assert ulps(O.math_probe("two_atanh_wide", z), exact).max() <= 1


def test_high_word_threshold_of_the_wide_rows():
    """rows outside the far / middle tiers test z^2 < 0x3FC47AE1'00000000 (just below 0.16) in the row kernels"""
    hi = lambda x: int(np.float64(x).view(np.uint64) >> np.uint64(32))
    assert hi(0.16) == 0x3FC47AE1 and hi(0.0025) == 0x3F647AE1
    tw = np.uint64(0x3FC47AE100000000).view(np.float64)
    tf = np.uint64(0x3F647AE100000000).view(np.float64)
    assert 0.16 - 4e-8 < tw < 0.16 and 0.0025 - 6e-10 < tf < 0.0025
    assert tw < 0.1601 and tf < 0.0025      # inside the ranges the polynomials were fitted on


def test_sincos():
    rng = np.random.default_rng(2)
    th = np.concatenate([rng.uniform(-100, 100, 400000), rng.uniform(-1e4, 1e4, 100000), [0.0, np.pi / 4, -np.pi / 2]])
    s, c = O.math_probe("sincos", th)
    assert np.abs(s - np.sin(th)).max() < 3e-16 and np.abs(c - np.cos(th)).max() < 3e-16
    assert np.abs(s * s + c * c - 1).max() < 5e-16


def test_roots():
    rng = np.random.default_rng(3)
    x = 10.0 ** rng.uniform(-30, 30, 200000)
    assert ulps(O.math_probe("root8", x), x ** 0.125).max() <= 2
    assert ulps(O.math_probe("root4", x), x ** 0.25).max() <= 2


def test_high_word_thresholds_of_the_range_tests():
    """The pair sum tests `z^2 < 0x3FA47AE1'00000000` and `q^2 < 0x3F847AE1'00000000` on the high words: the constants
    are the high words of 0.04 and 0.01, the thresholds sit just below them (inside the range the polynomials were
    fitted on), and for non-negative doubles the high-word order is the value order."""
    hi = lambda x: int(np.float64(x).view(np.uint64) >> np.uint64(32))
    assert hi(0.04) == 0x3FA47AE1 and hi(0.01) == 0x3F847AE1
    tz = np.uint64(0x3FA47AE100000000).view(np.float64)
    tq = np.uint64(0x3F847AE100000000).view(np.float64)
    assert 0.04 - 1e-8 < tz < 0.04 and 0.01 - 3e-9 < tq < 0.01
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.uniform(0, 0.1, 100000), tz + np.arange(-50, 50) * np.spacing(tz), [0.0, np.inf, np.nan]])
    hw = (x.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    with np.errstate(invalid="ignore"):
        assert np.array_equal(hw < 0x3FA47AE1, x < tz)          # NaN and inf fail both
