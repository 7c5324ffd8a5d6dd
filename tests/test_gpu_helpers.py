"""Device math helpers through the C ABI (one thin kernel each) against the reference's known answers, the
reference-generated golden vectors and the CPU oracle.  Tolerances: 1e-6 is the bar BASELINE.json sets for the
Fresnel / refraction functions; the device actually agrees to ~1e-15 and the tests hold it to 1e-12."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pvt_oracle
from pvtrace_b200.engine import _cuda

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def dev_fresnel(angle, n1, n2):
    angle, n1, n2 = _f64(angle), _f64(n1), _f64(n2)
    out = np.zeros_like(angle)
    _cuda.check(_cuda.load_library().pvt_test_fresnel_reflectivity(len(angle), _cuda._vp(angle), _cuda._vp(n1),
                                                                   _cuda._vp(n2), _cuda._vp(out), 0), "fresnel")
    return out


def dev_reflect(d, nrm):
    d, nrm = _f64(d), _f64(nrm)
    out = np.zeros_like(d)
    _cuda.check(_cuda.load_library().pvt_test_specular_reflect(len(d), _cuda._vp(d), _cuda._vp(nrm), _cuda._vp(out), 0), "reflect")
    return out


def dev_refract(d, nrm, n1, n2):
    d, nrm, n1, n2 = _f64(d), _f64(nrm), _f64(n1), _f64(n2)
    out = np.zeros_like(d)
    _cuda.check(_cuda.load_library().pvt_test_fresnel_refract(len(d), _cuda._vp(d), _cuda._vp(nrm), _cuda._vp(n1),
                                                              _cuda._vp(n2), _cuda._vp(out), 0), "refract")
    return out


def dev_intersect(gtype, params, o, d):
    g = np.ascontiguousarray(gtype, dtype=np.int32)
    params, o, d = _f64(params), _f64(o), _f64(d)
    nhit, ts = np.zeros(len(g), dtype=np.int32), np.zeros((len(g), 4))
    _cuda.check(_cuda.load_library().pvt_test_intersect(len(g), _cuda._vp(g), _cuda._vp(params), _cuda._vp(o), _cuda._vp(d),
                                                        _cuda._vp(nhit), _cuda._vp(ts), 0), "intersect")
    return nhit, ts


def dev_normal(gtype, params, p):
    g = np.ascontiguousarray(gtype, dtype=np.int32)
    params, p = _f64(params), _f64(p)
    out = np.zeros_like(p)
    _cuda.check(_cuda.load_library().pvt_test_local_normal(len(g), _cuda._vp(g), _cuda._vp(params), _cuda._vp(p), _cuda._vp(out), 0), "normal")
    return out


def dev_interp(x, xs, ys):
    x, xs, ys = _f64(x), _f64(xs), _f64(ys)
    out = np.zeros_like(x)
    _cuda.check(_cuda.load_library().pvt_test_interp(len(x), _cuda._vp(x), len(xs), _cuda._vp(xs), _cuda._vp(ys), _cuda._vp(out), 0), "interp")
    return out


def test_fresnel_known_answers_and_golden(gpu):
    assert abs(dev_fresnel([0.0], [1.0], [1.5])[0] - 0.04) < 1e-6          # tests/test_frensel_reflection.py:9-11
    g = np.load(os.path.join(GOLDEN, "optics.npz"))
    got = dev_fresnel(g["angle"], g["n1"], g["n2"])
    np.testing.assert_allclose(got, g["R"], rtol=0, atol=1e-6)               # the stated bar
    # away from the total-internal-reflection threshold (where R jumps to 1) the agreement is to rounding
    crit = np.where(g["n2"] < g["n1"], np.arcsin(np.minimum(g["n2"] / g["n1"], 1.0)), np.inf)
    smooth = np.abs(g["angle"] - crit) > 1e-6
    np.testing.assert_allclose(got[smooth], g["R"][smooth], rtol=0, atol=1e-12)
    np.testing.assert_allclose(got, pvt_oracle.fresnel_reflectivity(g["angle"], g["n1"], g["n2"]), rtol=0, atol=1e-6)


def test_reflect_refract_known_answers_and_golden(gpu):
    for normal in ((0.0, 0.0, 1.0), (0.0, 0.0, -1.0)):                       # test_frensel_reflection.py:13-27
        assert np.allclose(dev_reflect([(0.0, 0.0, -1.0)], [normal])[0], (0.0, 0.0, 1.0))
        assert np.allclose(dev_refract([(0.0, 0.0, -1.0)], [normal], [1.0], [1.5])[0], (0.0, 0.0, -1.0))
    g = np.load(os.path.join(GOLDEN, "optics.npz"))
    np.testing.assert_allclose(dev_reflect(g["d"], g["nrm"]), g["refl"], rtol=0, atol=1e-14)
    ok = g["refr_ok"]
    got = dev_refract(g["d"][ok], g["nrm"][ok], g["n1"][ok], g["n2"][ok])
    np.testing.assert_allclose(got, g["refr"][ok], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got, g["refr"][ok], rtol=0, atol=1e-12)


def test_intersectors_match_oracle(gpu):
    rng = np.random.default_rng(8)
    n = 300000
    gtype = rng.integers(0, 3, n).astype(np.int32)
    params = np.zeros((n, 4))
    params[:, :3] = rng.uniform(0.2, 3.0, size=(n, 3))
    o = rng.uniform(-4, 4, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:1000, 0] = 0.0  # axis-parallel rays exercise the slab guards
    d[:1000] /= np.linalg.norm(d[:1000], axis=1, keepdims=True)
    d[1000:1500] = (0.0, 0.0, 1.0)
    nhit, ts = dev_intersect(gtype, params, o, d)
    want_n, want_t = pvt_oracle.intersect(gtype, params, o, d)
    same = nhit == want_n
    assert same.mean() > 0.99999
    np.testing.assert_allclose(ts[same], want_t[same], rtol=1e-10, atol=1e-12)
    assert (nhit > 0).mean() > 0.05


def test_cylinder_literals_on_device(gpu):
    """tests/test_geometry_utils.py:63-101"""
    def hits(origin, direction):
        direction = np.asarray(direction, float) / np.linalg.norm(direction)
        nhit, ts = dev_intersect([2], [[1.0, 1.0, 0, 0]], [origin], [direction])
        return np.asarray(origin) + np.sort(ts[0, :nhit[0]])[:, None] * direction
    np.testing.assert_allclose(hits((0.2, 0.2, -1), (0, 0, 1)), ((0.2, 0.2, -0.5), (0.2, 0.2, 0.5)))
    np.testing.assert_allclose(hits((-2, 0.2, 0.0), (1.0, 0.2, -0.2)),
                               ((-0.9082895433880116, 0.41834209132239775, -0.2183420913223977), (0.5, 0.7, -0.5)))
    np.testing.assert_allclose(hits((0.0, 0.0, -1.5), (0.0, 1.0, 1.0))[0], (0.0, 1.0, -0.5))


def test_normals_match_oracle_and_golden(gpu):
    g = np.load(os.path.join(GOLDEN, "geometry.npz"))
    k = len(g["cyl_surf"])
    got = dev_normal(np.full(k, 2), np.tile([float(g["cyl_len"]), float(g["cyl_rad"]), 0, 0], (k, 1)), g["cyl_surf"])
    np.testing.assert_allclose(got, g["cyl_nrm"], atol=1e-9)
    rng = np.random.default_rng(9)
    n = 100000
    gtype = rng.integers(0, 3, n).astype(np.int32)
    params = np.zeros((n, 4))
    params[:, :3] = rng.uniform(0.2, 3.0, size=(n, 3))
    p = rng.uniform(-1.5, 1.5, size=(n, 3))
    np.testing.assert_allclose(dev_normal(gtype, params, p), pvt_oracle.local_normal(gtype, params, p), atol=1e-12)


def test_interp_bisection_and_hinted_paths(gpu):
    g = np.load(os.path.join(GOLDEN, "distribution.npz"))
    xq = np.concatenate([g["xq"], [0.0, 300.0, 1000.0, 2000.0], g["x"][:20]])
    # uniform grid -> guessed-bracket path; values are the reference Distribution's
    np.testing.assert_allclose(dev_interp(g["xq"], g["x"], g["y"]), g["value"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(dev_interp(xq, g["x"], g["y"]), pvt_oracle.interp(xq, g["x"], g["y"]), rtol=1e-13, atol=1e-300)
    # non-uniform abscissa (a CDF) -> bisection path
    np.testing.assert_allclose(dev_interp(g["pq"], g["cdf"], g["x"]), g["sample"], rtol=1e-12)
    xs = np.sort(np.random.default_rng(4).uniform(0, 1, 57))
    ys = np.random.default_rng(5).uniform(0, 1, 57)
    q = np.random.default_rng(6).uniform(-0.1, 1.1, 5000)
    np.testing.assert_allclose(dev_interp(q, xs, ys), pvt_oracle.interp(q, xs, ys), rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(dev_interp(q, [0.5], [7.0]), 7.0)  # single-knot (constant) table


def test_rng_streams_bit_identical_to_oracle(gpu):
    lib = _cuda.load_library()
    for mode in (_cuda.RNG_PHILOX, _cuda.RNG_XOSHIRO):
        out = np.zeros((1000, 37))
        _cuda.check(lib.pvt_test_rng_uniform(1000, 37, 2 ** 63 + 12345, 77, mode, _cuda._vp(out), 0), "rng")
        want = pvt_oracle.rng_uniform(1000, 37, 2 ** 63 + 12345, 77, mode)
        assert (out == want).all()
        assert 0.0 <= out.min() and out.max() < 1.0 and abs(out.mean() - 0.5) < 0.01


def test_phase_functions_match_oracle(gpu):
    lib = _cuda.load_library()
    for ptype, prm in ((0, 0.0), (1, 0.6), (1, -0.3), (1, 0.0), (2, 0.5)):
        out = np.zeros((20000, 3))
        _cuda.check(lib.pvt_test_sample_phase(20000, ptype, C.c_double(prm), 5, _cuda.RNG_PHILOX, _cuda._vp(out), 0), "phase")
        want = pvt_oracle.sample_phase(20000, ptype, prm, 5, _cuda.RNG_PHILOX)
        np.testing.assert_allclose(out, want, rtol=0, atol=1e-9)
        np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-12)


def dev_math(op, a, b=None):
    a = _f64(a)
    b = None if b is None else _f64(b)
    out = np.zeros_like(a)
    _cuda.check(_cuda.load_library().pvt_test_math(len(a), op, _cuda._vp(a), _cuda._vp(b), _cuda._vp(out), 0), "math")
    return out


def test_lean_arithmetic_against_libm(gpu):
    """The trace kernels' own log(1 - u), a / b, 1 / x and sqrt (pvt_math.cuh: no special cases, no slow paths) against
    numpy's: the logarithm to one ulp (the library log() it replaces promises the same), the division and the
    reciprocal correctly rounded in all but a 1e-9 sliver of the arguments (and never more than one ulp off), the
    square root of [0, 1] arguments equal to the IEEE one."""
    rng = np.random.default_rng(11)
    u = rng.integers(0, 1 << 53, 2_000_000, dtype=np.uint64).astype(np.float64) / float(1 << 53)
    x = np.concatenate([1.0 - u, np.ldexp(1.0 - u[:200000], -rng.integers(0, 53, 200000)),
                        [1.0, 0.5, 2.0 ** -53, np.nextafter(1.0, 0.0), np.sqrt(0.5), np.sqrt(2.0) / 2 * (1 + 1e-16)]])
    x = x[x > 0]
    got, ref = dev_math(0, x), np.log(x)
    ulp = np.spacing(np.abs(ref))
    err = np.abs(got - ref) / np.where(ulp > 0, ulp, 1.0)
    assert err.max() <= 1.0 + 1e-9, err.max()
    assert got[x == 1.0].tolist() == [0.0] * int((x == 1.0).sum())
    assert (got != ref).mean() < 0.15  # mostly the correctly rounded value
    # a / b over the operand ranges of the path: indices, absorption coefficients, Fresnel terms, knot spacings
    a = rng.uniform(-1e3, 1e3, 2_000_000) * 10.0 ** rng.integers(-8, 4, 2_000_000)
    b = rng.uniform(0.5, 2.0, 2_000_000) * 10.0 ** rng.integers(-8, 5, 2_000_000) * rng.choice([-1.0, 1.0], 2_000_000)
    q, exact = dev_math(1, a, b), a / b
    off = q != exact
    assert off.mean() < 1e-6, off.mean()
    assert np.all(np.abs(q[off] - exact[off]) <= np.spacing(np.abs(exact[off])))
    d = np.concatenate([rng.uniform(-1.0, 1.0, 2_000_000), 10.0 ** rng.uniform(-300, 0, 200000), [1.0, -1.0, 1e-300]])
    d = d[np.abs(d) >= 1e-300]
    r, exact = dev_math(2, d), 1.0 / d
    off = r != exact
    assert off.mean() < 1e-6, off.mean()
    assert np.all(np.abs(r[off] - exact[off]) <= np.spacing(np.abs(exact[off])))
    # square root of [0, 1] arguments (1 - c^2): the IEEE value
    x = np.concatenate([rng.uniform(0, 1, 2_000_000), 1.0 - rng.uniform(0, 1, 500000) ** 2, [0.0, 1.0, 1e-300, 2.0 ** -52]])
    np.testing.assert_array_equal(dev_math(3, x), np.sqrt(x))
