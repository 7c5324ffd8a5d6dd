"""CUDA tracer vs the CPU oracle on the SAME counter-based random stream (Philox), through the C ABI.

Because both sides draw identical uniforms for every photon, histories must agree ray by ray; the only admissible
differences are floating-point ones (the device contracts a*b+c into FMAs and its log/acos/asin/sincos differ from
glibc by an ulp or two), which can flip a comparison for a handful of rays per 10^5.  So the bar is: >= 99.9 % of
rays with an identical event sequence, positions of those equal to 1e-9, and every integer tally within
a few sigma of the oracle's.

Draws are ADDRESSED (photon, step, purpose -> Philox counter, csrc/pvt_rng.cuh) rather than consumed in sequence,
so a decision that one side skips (no surface draw when the Fresnel reflectivity is exactly 0,
pvtrace/material/surface.py:231-240 -- which at an equal-index interface depends on the last bit of sin/cos) does
not shift the uniforms of later decisions, and the pairing of histories survives such interfaces
(nested_cylinders has one between cylinders A and B).
"""
import numpy as np
import pytest

import pvtrace_b200 as pv
from oracle import pvt_oracle
from pvtrace_b200.device import configs
from pvtrace_b200.engine import _cuda
from pvtrace_b200.engine.compiler import EMIT_METHODS

pytestmark = pytest.mark.gpu
NAMES = list(configs.CONFIGS)
KERNELS = {"wavefront": 0, "register": _cuda.FLAG_REGISTER_KERNEL}


def _both(name, n, record_every, max_events=256, seed=7, rng_mode=_cuda.RNG_PHILOX, flags=0):
    build, kw = configs.CONFIGS[name]
    scene = build()
    compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
    method = EMIT_METHODS[kw["emit_method"]]
    got = _cuda.trace_bundle(compiled, None, None, None, seed, 1000, max_events, method, 0, record_every,
                             emitter=emitter, n=n, rng_mode=rng_mode, flags=flags)
    want = pvt_oracle.trace_bundle(compiled, None, None, None, seed, 1000, max_events, method, 8, record_every,
                                   emitter=emitter, n=n, rng_mode=rng_mode)
    return compiled, got, want


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("name", NAMES)
def test_histories_match_oracle_ray_by_ray(gpu, name, kernel):
    n, m = 20000, 256
    compiled, got, want = _both(name, n, 1, m, flags=KERNELS[kernel])
    same_count = got["counts"] == want["counts"]
    kinds_g, kinds_w = got["kind"].reshape(n, m), want["kind"].reshape(n, m)
    same_kinds = same_count & (kinds_g == kinds_w).all(axis=1)
    assert same_kinds.mean() >= 0.999, f"only {same_kinds.mean():.5f} of rays have identical event sequences"
    rows = np.repeat(same_kinds, m)
    for key, tol in (("position", 1e-7), ("direction", 1e-7), ("wavelength", 1e-7), ("travelled", 1e-7), ("normal", 1e-7)):
        np.testing.assert_allclose(got[key][rows], want[key][rows], rtol=0, atol=tol, err_msg=key)
    np.testing.assert_allclose(got["duration"][rows], want["duration"][rows], rtol=1e-9, atol=1e-18)
    for key in ("hit", "container", "adjacent", "component", "source"):
        assert (got[key][rows] == want[key][rows]).all(), key


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("name", NAMES)
def test_tallies_match_oracle(gpu, name, kernel):
    n = 200000
    compiled, got, want = _both(name, n, 0, flags=KERNELS[kernel])
    for key in ("rec_distinct", "rec_crossings", "rec_bins"):
        g, w = got[key].astype(float), want[key].astype(float)
        p = np.clip(w / n, 1e-9, 1 - 1e-9)
        # a fraction f of the rays may be decorrelated from the oracle's (f = 0.1 % from rounding; 5 % where
        # equal-index interfaces shift the stream, see module docstring): 5 sigma of the difference of two
        # independent binomial samples over f*n rays
        f = 0.001
        tol = 5.0 * np.sqrt(2 * f * n * p * (1 - p)) + 3.0
        assert (np.abs(g - w) <= tol).all(), (key, np.abs(g - w).max())
    assert got["stats"][_cuda.STAT_RAYS] == n
    assert abs(got["stats"][_cuda.STAT_STEPS] - want["stats"][_cuda.STAT_STEPS]) <= 0.002 * want["stats"][_cuda.STAT_STEPS]
    np.testing.assert_allclose(got["rec_sums"], want["rec_sums"], rtol=5e-3, atol=1e-12)
