"""ctypes binding of the CPU oracle (oracle/libpvt_oracle.so, built by `make -C oracle oracle`).

TEST INFRASTRUCTURE.  Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and the `cpu_baseline` /
`--impl reference` legs of bench.py may import this module; the product package never does.  The struct layouts
are the C ABI's (include/pvtrace_b200.h) and are shared with the product's ctypes binding so that the oracle and
the CUDA library are fed byte-identical tables.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from pvtrace_b200.engine import _cuda as abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpvt_oracle.so")
_lib = None


def build(force=False):
    """Compile the oracle (and, when the reference tree is present, the reference's own kernel into oracle/_ref)."""
    src = os.path.join(HERE, "pvt_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def build_ref():
    from oracle import ref_loader

    if ref_loader.ref_kernel_path() is None and ref_loader.reference_available():
        subprocess.run(["make", "-C", HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
    return ref_loader.ref_kernel_path()


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        lib.pvt_oracle_trace_bundle.restype = C.c_int
        lib.pvt_oracle_trace_bundle.argtypes = [C.POINTER(abi.PvtScene), C.POINTER(abi.PvtEmit), vp, vp, vp,
                                                C.POINTER(abi.PvtParams), C.POINTER(abi.PvtOut), C.c_int]
        lib.pvt_oracle_emit_bundle.restype = C.c_int
        lib.pvt_oracle_emit_bundle.argtypes = [C.POINTER(abi.PvtEmit), vp, vp, vp, C.c_int64, C.c_int64, C.c_uint64]
        lib.pvt_oracle_intersect_bundle.restype = C.c_int
        lib.pvt_oracle_intersect_bundle.argtypes = [C.POINTER(abi.PvtScene), vp, vp, C.c_int64, vp, vp, vp, vp]
        _lib = lib
    return _lib


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def trace_bundle(compiled, positions, directions, wavelengths, seed, maxsteps, max_events, emit_method,
                 num_threads=1, record_every=1, *, emitter=None, n=None, first_index=0, rng_mode=abi.RNG_PHILOX):
    """Same signature and result dict as the product's `_cuda.trace_bundle`, computed on the CPU."""
    lib = load()
    if positions is not None:
        positions, directions, wavelengths = _f64(positions), _f64(directions), _f64(wavelengths)
        n = positions.shape[0]
    scene, keep = abi.marshal_scene(compiled)
    emit_struct = None
    if emitter is not None:
        emit_struct, keep_e = abi.marshal_emitter(emitter)
        keep += keep_e
    params = abi.make_params(n, seed, maxsteps, max_events, emit_method, record_every, first_index, rng_mode, 0)
    data, out = abi.allocate_outputs(compiled, n, max_events, record_every)
    status = lib.pvt_oracle_trace_bundle(C.byref(scene), C.byref(emit_struct) if emit_struct is not None else None,
                                         _vp(positions), _vp(directions), _vp(wavelengths), C.byref(params),
                                         C.byref(out), int(num_threads))
    if status != 0:
        raise RuntimeError(f"oracle trace failed ({status})")
    return abi.finalize_outputs(compiled, data, n, record_every)


def emit_bundle(emitter, n, seed, first_index=0):
    lib = load()
    struct, keep = abi.marshal_emitter(emitter)
    pos, direction, wl = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
    status = lib.pvt_oracle_emit_bundle(C.byref(struct), _vp(pos), _vp(direction), _vp(wl), int(n), int(first_index),
                                        int(seed) & 0xFFFFFFFFFFFFFFFF)
    if status != 0:
        raise RuntimeError("oracle emit failed")
    return pos, direction, wl


def intersect_bundle(compiled, positions, directions):
    lib = load()
    positions, directions = _f64(positions), _f64(directions)
    n = positions.shape[0]
    scene, keep = abi.marshal_scene(compiled)
    t0 = np.zeros(n)
    hit, container, adjacent = (np.zeros(n, dtype=np.int32) for _ in range(3))
    lib.pvt_oracle_intersect_bundle(C.byref(scene), _vp(positions), _vp(directions), n, _vp(t0), _vp(hit),
                                    _vp(container), _vp(adjacent))
    return t0, hit, container, adjacent


def surface_event(compiled, hit, container, adjacent, positions, directions, wavelengths, p1=None, p2=None):
    """One surface interaction per row (points ON the surface of node `hit`): reflectivity, reflected direction and
    transmitted direction (NaN where the reflectivity is 1) -- the facet / coating extension on its own."""
    lib = load()
    positions, directions, wavelengths = _f64(positions), _f64(directions), _f64(wavelengths)
    n = len(wavelengths)
    ids = [np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.int32), (n,))) for v in (hit, container, adjacent)]
    p1 = None if p1 is None else _f64(p1)
    p2 = None if p2 is None else _f64(p2)
    scene, keep = abi.marshal_scene(compiled)
    refl, r_dir, t_dir = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3))
    vp = C.c_void_p
    fn = lib.pvt_oracle_surface_event
    fn.restype, fn.argtypes = C.c_int, [C.POINTER(abi.PvtScene), C.c_int64] + [vp] * 11
    if fn(C.byref(scene), n, _vp(ids[0]), _vp(ids[1]), _vp(ids[2]), _vp(positions), _vp(directions), _vp(wavelengths),
          _vp(p1), _vp(p2), _vp(refl), _vp(r_dir), _vp(t_dir)) != 0:
        raise RuntimeError("pvt_oracle_surface_event failed")
    return refl, r_dir, t_dir


def _call_helper(name, argtypes, *args):
    fn = getattr(load(), name)
    fn.restype, fn.argtypes = C.c_int, argtypes
    if fn(*args) != 0:
        raise RuntimeError(f"{name} failed")


def fresnel_reflectivity(angle, n1, n2):
    angle, n1, n2 = _f64(angle), _f64(n1), _f64(n2)
    out = np.zeros_like(angle)
    vp = C.c_void_p
    _call_helper("pvt_oracle_fresnel_reflectivity", [C.c_int64, vp, vp, vp, vp], len(angle), _vp(angle), _vp(n1),
                 _vp(n2), _vp(out))
    return out


def specular_reflect(d, nrm):
    d, nrm = _f64(d), _f64(nrm)
    out = np.zeros_like(d)
    vp = C.c_void_p
    _call_helper("pvt_oracle_specular_reflect", [C.c_int64, vp, vp, vp], len(d), _vp(d), _vp(nrm), _vp(out))
    return out


def fresnel_refract(d, nrm, n1, n2):
    d, nrm, n1, n2 = _f64(d), _f64(nrm), _f64(n1), _f64(n2)
    out = np.zeros_like(d)
    vp = C.c_void_p
    _call_helper("pvt_oracle_fresnel_refract", [C.c_int64, vp, vp, vp, vp, vp], len(d), _vp(d), _vp(nrm), _vp(n1),
                 _vp(n2), _vp(out))
    return out


def intersect(geom_type, params, o, d):
    g = np.ascontiguousarray(geom_type, dtype=np.int32)
    params, o, d = _f64(params), _f64(o), _f64(d)
    nhit, ts = np.zeros(len(g), dtype=np.int32), np.zeros((len(g), 4))
    vp = C.c_void_p
    _call_helper("pvt_oracle_intersect", [C.c_int64, vp, vp, vp, vp, vp, vp], len(g), _vp(g), _vp(params), _vp(o),
                 _vp(d), _vp(nhit), _vp(ts))
    return nhit, ts


def local_normal(geom_type, params, p):
    g = np.ascontiguousarray(geom_type, dtype=np.int32)
    params, p = _f64(params), _f64(p)
    out = np.zeros_like(p)
    vp = C.c_void_p
    _call_helper("pvt_oracle_local_normal", [C.c_int64, vp, vp, vp, vp], len(g), _vp(g), _vp(params), _vp(p), _vp(out))
    return out


def interp(x, xs, ys):
    x, xs, ys = _f64(x), _f64(xs), _f64(ys)
    out = np.zeros_like(x)
    vp = C.c_void_p
    _call_helper("pvt_oracle_interp", [C.c_int64, vp, C.c_int32, vp, vp, vp], len(x), _vp(x), len(xs), _vp(xs), _vp(ys),
                 _vp(out))
    return out


def rng_uniform(n_rays, n_draws, seed, first_index=0, rng_mode=abi.RNG_PHILOX):
    out = np.zeros((n_rays, n_draws))
    _call_helper("pvt_oracle_rng_uniform", [C.c_int64, C.c_int32, C.c_uint64, C.c_int64, C.c_int32, C.c_void_p],
                 n_rays, n_draws, int(seed) & 0xFFFFFFFFFFFFFFFF, first_index, rng_mode, _vp(out))
    return out


def sample_phase(n, phase_type, phase_param, seed, rng_mode=abi.RNG_PHILOX):
    out = np.zeros((n, 3))
    _call_helper("pvt_oracle_sample_phase", [C.c_int64, C.c_int32, C.c_double, C.c_uint64, C.c_int32, C.c_void_p], n,
                 phase_type, float(phase_param), int(seed) & 0xFFFFFFFFFFFFFFFF, rng_mode, _vp(out))
    return out


def philox4x32_10(counter, key):
    ctr = (C.c_uint32 * 4)(*counter)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    fn = load().pvt_oracle_philox4x32_10
    fn.restype = None
    fn(ctr, k, out)
    return tuple(out)
