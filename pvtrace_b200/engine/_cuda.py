"""ctypes binding of the C ABI in include/pvtrace_b200.h (libpvtrace_b200.so, built by csrc/build.py).

This is the reference-side stub a pvtrace maintainer would add in place of `from pvtrace.engine import _kernel`
(pvtrace/engine/api.py:215): `trace_bundle` below has the reference kernel's signature and returns the same
dict (pvtrace/engine/_kernel.pyx:903-1115).  There is no CPU fallback: if the library is missing or no CUDA
device is usable every call raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    "PVTRACE_B200_LIB", os.path.join(os.path.dirname(HERE), "csrc", "libpvtrace_b200.so"))

FLAG_REGISTER_KERNEL = 1  # PVT_FLAG_REGISTER_KERNEL: force the one-photon-per-lane kernel
RNG_PHILOX, RNG_XOSHIRO = 0, 1
RNG_MODES = {"philox": RNG_PHILOX, "xoshiro": RNG_XOSHIRO}
NSTATS = 32
STAT_STEPS, STAT_RAYS, STAT_LAUNCHES, STAT_EVENTS, STAT_H2D_BYTES = 0, 1, 2, 3, 4

_P_I32, _P_F64 = C.POINTER(C.c_int32), C.POINTER(C.c_double)
_P_I64, _P_U8 = C.POINTER(C.c_int64), C.POINTER(C.c_uint8)

_SCENE_COUNTS = ("n_nodes", "root_id", "n_components", "n_abs_knots", "n_ems_knots", "n_recorders", "n_hists",
                 "total_bins", "n_facets", "n_refl_knots")
# (struct field, CompiledScene attribute, dtype) in header order
_SCENE_TABLES = (
    ("geom_type", "geom_type", np.int32), ("geom_params", "geom_params", np.float64),
    ("local_to_world", "local_to_world", np.float64), ("world_to_local", "world_to_local", np.float64),
    ("refractive_index", "refractive_index", np.float64), ("surface_type", "surface_type", np.int32),
    ("comp_start", "comp_start", np.int32), ("comp_count", "comp_count", np.int32),
    ("comp_type", "comp_type", np.int32), ("comp_qy", "comp_qy", np.float64),
    ("comp_tau_rad", "comp_tau_rad", np.float64), ("comp_tau_nr", "comp_tau_nr", np.float64),
    ("comp_phase_type", "comp_phase_type", np.int32), ("comp_phase_param", "comp_phase_param", np.float64),
    ("comp_abs_start", "comp_abs_start", np.int32), ("comp_abs_n", "comp_abs_n", np.int32),
    ("comp_ems_start", "comp_ems_start", np.int32), ("comp_ems_n", "comp_ems_n", np.int32),
    ("abs_x", "abs_x", np.float64), ("abs_y", "abs_y", np.float64),
    ("ems_x", "ems_x", np.float64), ("ems_cdf", "ems_cdf", np.float64),
    ("rec_node", "rec_node", np.int32), ("rec_event", "rec_event", np.int32),
    ("rec_has_facet", "rec_has_facet", np.int32), ("rec_facet", "rec_facet", np.float64),
    ("rec_atol", "rec_atol", np.float64), ("rec_hist_start", "rec_hist_start", np.int32),
    ("rec_hist_n", "rec_hist_n", np.int32), ("hist_prop_a", "hist_prop_a", np.int32),
    ("hist_prop_b", "hist_prop_b", np.int32), ("hist_na", "hist_na", np.int32), ("hist_nb", "hist_nb", np.int32),
    ("hist_lo_a", "hist_lo_a", np.float64), ("hist_hi_a", "hist_hi_a", np.float64),
    ("hist_lo_b", "hist_lo_b", np.float64), ("hist_hi_b", "hist_hi_b", np.float64),
    ("hist_offset", "hist_offset", np.int32),
    ("facet_start", "facet_start", np.int32), ("facet_count", "facet_count", np.int32),
    ("facet_normal", "facet_normal", np.float64), ("facet_atol", "facet_atol", np.float64),
    ("facet_reflectivity", "facet_reflectivity", np.float64), ("facet_flags", "facet_flags", np.int32),
    ("facet_region", "facet_region", np.float64), ("facet_refl_start", "facet_refl_start", np.int32),
    ("facet_refl_n", "facet_refl_n", np.int32), ("refl_x", "refl_x", np.float64), ("refl_y", "refl_y", np.float64),
)


def _ptr_type(dtype):
    return _P_I32 if dtype == np.int32 else _P_F64


class PvtScene(C.Structure):
    _fields_ = [(name, C.c_int32) for name in _SCENE_COUNTS] + [
        (field, _ptr_type(dtype)) for field, _, dtype in _SCENE_TABLES]


_EMIT_TABLES = (
    ("light_to_world", np.float64), ("pos_kind", np.int32), ("pos_param", np.float64), ("dir_kind", np.int32),
    ("dir_param", np.float64), ("wl_kind", np.int32), ("wl_param", np.float64), ("wl_start", np.int32),
    ("wl_n", np.int32), ("wl_x", np.float64), ("wl_cdf", np.float64),
)


class PvtEmit(C.Structure):
    _fields_ = [("n_lights", C.c_int32), ("n_wl_knots", C.c_int32)] + [
        (field, _ptr_type(dtype)) for field, dtype in _EMIT_TABLES]


class PvtParams(C.Structure):
    _fields_ = [("n", C.c_int64), ("first_index", C.c_int64), ("seed", C.c_uint64), ("record_every", C.c_int64),
                ("maxsteps", C.c_int32), ("max_events", C.c_int32), ("emit_method", C.c_int32),
                ("rng_mode", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32)]


_OUT_FIELDS = (
    ("counts", np.int32), ("rec_distinct", np.int64), ("rec_crossings", np.int64), ("rec_sums", np.float64),
    ("rec_bins", np.int64), ("kind", np.uint8), ("hit", np.int32), ("container", np.int32),
    ("adjacent", np.int32), ("component", np.int32), ("source", np.int32), ("position", np.float64),
    ("direction", np.float64), ("normal", np.float64), ("wavelength", np.float64), ("travelled", np.float64),
    ("duration", np.float64), ("stats", np.int64),
)
_CTYPE = {np.int32: _P_I32, np.int64: _P_I64, np.float64: _P_F64, np.uint8: _P_U8}


class PvtOut(C.Structure):
    _fields_ = [(field, _CTYPE[dtype]) for field, dtype in _OUT_FIELDS]


def _as_ptr(array, dtype):
    return array.ctypes.data_as(_CTYPE[dtype])


def marshal_scene(compiled):
    """PvtScene for a CompiledScene-like object.  Returns (struct, keepalive list).  Tables are immutable once
    compiled, so the struct is built once per object (repeated bundles of one scene: ~0.1 ms per call otherwise)."""
    cached = getattr(compiled, "_pvt_marshalled_scene", None)
    if cached is not None:
        return cached[0], list(cached[1])
    s, keep = PvtScene(), []
    for field, attr, dtype in _SCENE_TABLES:
        value = getattr(compiled, attr, None)
        if value is None:  # reference CompiledScene objects have no facet tables, older ones no coatings
            n_facets = int(getattr(compiled, "n_facets", 0))
            if field in ("facet_start", "facet_count"):
                value = np.zeros(len(compiled.geom_type), dtype=np.int32)
            elif field == "facet_region":
                value = np.tile(np.array([-np.inf] * 3 + [np.inf] * 3), (n_facets, 1))
            elif field in ("facet_refl_start", "facet_refl_n"):
                value = np.zeros(n_facets, dtype=np.int32)
            else:
                value = np.zeros(0, dtype=dtype)
        arr = np.ascontiguousarray(value, dtype=dtype)
        keep.append(arr)
        setattr(s, field, _as_ptr(arr, dtype))
    s.n_nodes = len(compiled.geom_type)
    s.root_id = int(compiled.root_id)
    s.n_components = len(compiled.comp_type)
    s.n_abs_knots = len(compiled.abs_x)
    s.n_ems_knots = len(compiled.ems_x)
    s.n_recorders = len(compiled.rec_node)
    s.n_hists = len(compiled.hist_offset)
    s.total_bins = int(compiled.total_bins)
    s.n_facets = int(getattr(compiled, "n_facets", 0))
    s.n_refl_knots = len(getattr(compiled, "refl_x", ()))
    try:
        compiled._pvt_marshalled_scene = (s, tuple(keep))
    except AttributeError:  # objects without a __dict__ (foreign table holders) are marshalled every time
        pass
    return s, keep


def marshal_emitter(emitter):
    e, keep = PvtEmit(), []
    for field, dtype in _EMIT_TABLES:
        arr = np.ascontiguousarray(getattr(emitter, field), dtype=dtype)
        keep.append(arr)
        setattr(e, field, _as_ptr(arr, dtype))
    e.n_lights = int(emitter.n_lights)
    e.n_wl_knots = len(emitter.wl_x)
    return e, keep


def make_params(n, seed, maxsteps, max_events, emit_method, record_every, first_index=0, rng_mode=RNG_PHILOX,
                device=0, flags=0):
    p = PvtParams()
    p.n, p.first_index, p.seed = int(n), int(first_index), int(seed) & 0xFFFFFFFFFFFFFFFF
    p.record_every, p.maxsteps, p.max_events = int(record_every), int(maxsteps), int(max_events)
    p.emit_method, p.rng_mode, p.device, p.flags = int(emit_method), int(rng_mode), int(device), int(flags)
    return p


def allocate_outputs(compiled, n, max_events, record_every):
    """Output arrays of one bundle, laid out and initialised like the reference (_kernel.pyx:1035-1047).
    Returns (dict of arrays, PvtOut)."""
    n_recorded = (n + record_every - 1) // record_every if record_every > 0 else 0
    rows = n_recorded * max_events
    n_rec = len(compiled.rec_node)
    data = {
        "counts": np.zeros(max(n_recorded, 1), dtype=np.int32),
        "rec_distinct": np.zeros(max(n_rec, 1), dtype=np.int64),
        "rec_crossings": np.zeros(max(n_rec, 1), dtype=np.int64),
        "rec_sums": np.zeros(max(n_rec, 1) * 8, dtype=np.float64),
        "rec_bins": np.zeros(max(int(compiled.total_bins), 1), dtype=np.int64),
        "kind": np.zeros(rows, dtype=np.uint8),
        "hit": np.full(rows, -1, dtype=np.int32), "container": np.full(rows, -1, dtype=np.int32),
        "adjacent": np.full(rows, -1, dtype=np.int32), "component": np.full(rows, -1, dtype=np.int32),
        "source": np.full(rows, -1, dtype=np.int32),
        "position": np.zeros((rows, 3)), "direction": np.zeros((rows, 3)), "normal": np.zeros((rows, 3)),
        "wavelength": np.zeros(rows), "travelled": np.zeros(rows), "duration": np.zeros(rows),
        "stats": np.zeros(NSTATS, dtype=np.int64),
    }
    out = PvtOut()
    for field, dtype in _OUT_FIELDS:
        setattr(out, field, _as_ptr(data[field], dtype))
    return data, out


def finalize_outputs(compiled, data, n, record_every):
    """Trim the padded arrays to the reference's shapes (_kernel.pyx:1097-1115)."""
    n_recorded = (n + record_every - 1) // record_every if record_every > 0 else 0
    n_rec = len(compiled.rec_node)
    data["counts"] = data["counts"][:n_recorded]
    data["rec_distinct"] = data["rec_distinct"][:n_rec]
    data["rec_crossings"] = data["rec_crossings"][:n_rec]
    data["rec_sums"] = data["rec_sums"][: n_rec * 8].reshape(n_rec, 4, 2)
    data["rec_bins"] = data["rec_bins"][: int(compiled.total_bins)]
    return data


class LibraryError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen the CUDA library and declare prototypes.  Raises LibraryError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryError(
            f"{LIB_PATH} not found: build it with `python -m pvtrace_b200.csrc.build` (needs nvcc). "
            "pvtrace_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    protos = {
        "pvt_version": (C.c_int, []),
        "pvt_device_count": (C.c_int, []),
        "pvt_last_error": (C.c_char_p, []),
        "pvt_struct_sizes": (None, [C.POINTER(C.c_int32)]),
        "pvt_measure_fp64_peak": (C.c_int, [C.c_int, _P_F64]),
        "pvt_trace_bundle": (C.c_int, [C.POINTER(PvtScene), C.POINTER(PvtEmit), vp, vp, vp,
                                       C.POINTER(PvtParams), C.POINTER(PvtOut), _P_F64]),
        "pvt_trace_bundle_devices": (C.c_int, [C.POINTER(PvtScene), C.POINTER(PvtEmit), vp, vp, vp,
                                               C.POINTER(PvtParams), C.c_int32, _P_I32, C.POINTER(PvtOut), _P_F64]),
        "pvt_context_create": (C.c_int, [C.POINTER(PvtScene), C.POINTER(PvtEmit), C.c_int, C.POINTER(vp)]),
        "pvt_context_destroy": (C.c_int, [vp]),
        "pvt_context_reset": (C.c_int, [vp, vp]),
        "pvt_trace_device": (C.c_int, [vp, vp, vp, vp, C.POINTER(PvtParams), vp]),
        "pvt_context_read": (C.c_int, [vp, C.POINTER(PvtOut), vp]),
        "pvt_context_pack_tallies": (C.c_int, [vp, C.POINTER(vp), _P_I64, vp]),
        "pvt_context_unpack_tallies": (C.c_int, [vp, vp]),
        "pvt_emit_device": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int64, C.c_uint64, vp]),
        "pvt_emit_bundle": (C.c_int, [C.POINTER(PvtEmit), vp, vp, vp, C.c_int64, C.c_int64, C.c_uint64, C.c_int]),
        "pvt_intersect_bundle": (C.c_int, [C.POINTER(PvtScene), vp, vp, C.c_int64, vp, vp, vp, vp, C.c_int, _P_F64]),
        "pvt_intersect_device": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp]),
        "pvt_intersect_device_packed": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp, vp]),
        "pvt_test_fresnel_reflectivity": (C.c_int, [C.c_int64, vp, vp, vp, vp, C.c_int]),
        "pvt_test_specular_reflect": (C.c_int, [C.c_int64, vp, vp, vp, C.c_int]),
        "pvt_test_fresnel_refract": (C.c_int, [C.c_int64, vp, vp, vp, vp, vp, C.c_int]),
        "pvt_test_intersect": (C.c_int, [C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int]),
        "pvt_test_local_normal": (C.c_int, [C.c_int64, vp, vp, vp, vp, C.c_int]),
        "pvt_test_interp": (C.c_int, [C.c_int64, vp, C.c_int32, vp, vp, vp, C.c_int]),
        "pvt_test_math": (C.c_int, [C.c_int64, C.c_int32, vp, vp, vp, C.c_int]),
        "pvt_test_rng_uniform": (C.c_int, [C.c_int64, C.c_int32, C.c_uint64, C.c_int64, C.c_int32, vp, C.c_int]),
        "pvt_test_sample_phase": (C.c_int, [C.c_int64, C.c_int32, C.c_double, C.c_uint64, C.c_int32, vp, C.c_int]),
    }
    for name, (restype, argtypes) in protos.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype, fn.argtypes = restype, argtypes
    sizes = (C.c_int32 * 4)()
    lib.pvt_struct_sizes(sizes)
    mine = [C.sizeof(PvtScene), C.sizeof(PvtEmit), C.sizeof(PvtParams), C.sizeof(PvtOut)]
    if list(sizes) != mine:
        raise LibraryError(f"struct layout mismatch between {LIB_PATH} {list(sizes)} and its ctypes mirror {mine}")
    _lib = lib
    return lib


EXPORTED_SYMBOLS = (
    "pvt_version", "pvt_device_count", "pvt_last_error", "pvt_struct_sizes", "pvt_measure_fp64_peak", "pvt_trace_bundle", "pvt_trace_bundle_devices", "pvt_context_create",
    "pvt_context_destroy", "pvt_context_reset", "pvt_trace_device", "pvt_context_read",
    "pvt_context_pack_tallies", "pvt_context_unpack_tallies", "pvt_emit_device", "pvt_emit_bundle",
    "pvt_intersect_bundle", "pvt_intersect_device", "pvt_intersect_device_packed", "pvt_test_fresnel_reflectivity",
    "pvt_test_specular_reflect", "pvt_test_fresnel_refract", "pvt_test_intersect", "pvt_test_local_normal",
    "pvt_test_interp", "pvt_test_rng_uniform", "pvt_test_sample_phase", "pvt_test_math",
)


def check(status, what):
    if status != 0:
        message = load_library().pvt_last_error()
        raise LibraryError(f"{what} failed ({status}): {message.decode() if message else 'unknown error'}")


def device_count():
    return int(load_library().pvt_device_count())


def measure_fp64_peak(device=0):
    """fp64 FMA throughput of the device in TFLOP/s, measured with the library's own micro-benchmark."""
    rate = C.c_double(0.0)
    check(load_library().pvt_measure_fp64_peak(int(device), C.byref(rate)), "pvt_measure_fp64_peak")
    return rate.value


def _vp(array):
    return None if array is None else array.ctypes.data_as(C.c_void_p)


def trace_bundle(compiled, positions, directions, wavelengths, seed, maxsteps, max_events, emit_method,
                 num_threads=0, record_every=1, *, emitter=None, n=None, first_index=0, rng_mode=RNG_PHILOX,
                 device=0, return_elapsed=False, flags=0, devices=None):
    """Drop-in for pvtrace.engine._kernel.trace_bundle (pvtrace/engine/_kernel.pyx:903-1115).

    `num_threads` is accepted for signature compatibility and ignored (the device schedules itself).  With an
    `emitter` (CompiledEmitter) the three ray arrays may be None and `n` rays are sampled on the device.
    `devices`: a sequence of CUDA ordinals to spread the bundle over (contiguous index slices, one host thread per
    device inside the library, tallies summed); None or one entry: the single `device`.
    """
    lib = load_library()
    if positions is not None:
        positions = np.ascontiguousarray(positions, dtype=np.float64)
        directions = np.ascontiguousarray(directions, dtype=np.float64)
        wavelengths = np.ascontiguousarray(wavelengths, dtype=np.float64)
        n = positions.shape[0]
    elif emitter is None or n is None:
        raise ValueError("either ray arrays or (emitter, n) are required")
    if len(compiled.geom_type) > 128:
        raise ValueError("Engine supports at most 128 geometry nodes.")
    scene, keep = marshal_scene(compiled)
    emit_struct = None
    if emitter is not None:
        emit_struct, keep_e = marshal_emitter(emitter)
        keep += keep_e
    params = make_params(n, seed, maxsteps, max_events, emit_method, record_every, first_index, rng_mode, device,
                         flags)
    data, out = allocate_outputs(compiled, n, max_events, record_every)
    elapsed = C.c_double(0.0)
    emit_ref = C.byref(emit_struct) if emit_struct is not None else None
    if devices is not None and len(devices) > 1:
        ids = np.ascontiguousarray(devices, dtype=np.int32)
        status = lib.pvt_trace_bundle_devices(C.byref(scene), emit_ref, _vp(positions), _vp(directions), _vp(wavelengths),
                                              C.byref(params), len(ids), _as_ptr(ids, np.int32), C.byref(out),
                                              C.byref(elapsed))
        check(status, "pvt_trace_bundle_devices")
    else:
        if devices is not None and len(devices) == 1:
            params.device = int(devices[0])
        status = lib.pvt_trace_bundle(C.byref(scene), emit_ref, _vp(positions), _vp(directions), _vp(wavelengths),
                                      C.byref(params), C.byref(out), C.byref(elapsed))
        check(status, "pvt_trace_bundle")
    data = finalize_outputs(compiled, data, n, record_every)
    if return_elapsed:
        return data, elapsed.value
    return data


class Context:
    """Resident scene on one device (pvt_context_*): tables uploaded once, tallies accumulate on the device,
    bundles are traced from device-resident ray arrays or from the on-device emitter.  Device pointers are
    plain integers (e.g. `tensor.data_ptr()`), streams are `torch.cuda.Stream.cuda_stream` integers or 0."""

    def __init__(self, compiled, emitter=None, device=0):
        self._lib = load_library()
        self.compiled, self.emitter, self.device = compiled, emitter, int(device)
        scene, keep = marshal_scene(compiled)
        emit_struct = None
        if emitter is not None:
            emit_struct, keep_e = marshal_emitter(emitter)
            keep += keep_e
        handle = C.c_void_p()
        status = self._lib.pvt_context_create(C.byref(scene), C.byref(emit_struct) if emit_struct is not None else None,
                                              self.device, C.byref(handle))
        check(status, "pvt_context_create")
        self._handle = handle

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.pvt_context_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self, stream=0):
        check(self._lib.pvt_context_reset(self._handle, C.c_void_p(stream)), "pvt_context_reset")

    def trace(self, n, seed, *, d_positions=0, d_directions=0, d_wavelengths=0, first_index=0, maxsteps=1000,
              max_events=128, emit_method=0, record_every=0, rng_mode=RNG_PHILOX, stream=0, flags=0):
        """Enqueue one bundle (asynchronous).  Null ray pointers => rays come from the context's emitter."""
        params = make_params(n, seed, maxsteps, max_events, emit_method, record_every, first_index, rng_mode,
                             self.device, flags)
        self._last = (int(n), int(max_events), int(record_every))
        status = self._lib.pvt_trace_device(self._handle, C.c_void_p(d_positions), C.c_void_p(d_directions),
                                            C.c_void_p(d_wavelengths), C.byref(params), C.c_void_p(stream))
        check(status, "pvt_trace_device")

    def read(self, stream=0):
        """Synchronise and return the reference-shaped result dict of everything accumulated since reset()
        (event log: the last trace only)."""
        n, max_events, record_every = getattr(self, "_last", (0, 1, 0))
        data, out = allocate_outputs(self.compiled, n, max_events, record_every)
        check(self._lib.pvt_context_read(self._handle, C.byref(out), C.c_void_p(stream)), "pvt_context_read")
        return finalize_outputs(self.compiled, data, n, record_every)

    def pack_tallies(self, stream=0):
        """(device pointer, count) of the packed float64 tally buffer to all-reduce across GPUs."""
        ptr, count = C.c_void_p(), C.c_int64()
        check(self._lib.pvt_context_pack_tallies(self._handle, C.byref(ptr), C.byref(count), C.c_void_p(stream)),
              "pvt_context_pack_tallies")
        return ptr.value or 0, int(count.value)

    def unpack_tallies(self, stream=0):
        check(self._lib.pvt_context_unpack_tallies(self._handle, C.c_void_p(stream)), "pvt_context_unpack_tallies")

    def emit(self, d_positions, d_directions, d_wavelengths, n, seed, first_index=0, stream=0):
        status = self._lib.pvt_emit_device(self._handle, C.c_void_p(d_positions), C.c_void_p(d_directions),
                                           C.c_void_p(d_wavelengths), int(n), int(first_index),
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, C.c_void_p(stream))
        check(status, "pvt_emit_device")

    def intersect_packed(self, d_positions, d_directions, n, d_t0, d_ids, stream=0):
        """The intersect stage with the ids of a ray in one uint32 (hit | container << 8 | adjacent << 16, 0xff = none)."""
        status = self._lib.pvt_intersect_device_packed(self._handle, C.c_void_p(d_positions), C.c_void_p(d_directions),
                                                       int(n), C.c_void_p(d_t0), C.c_void_p(d_ids), C.c_void_p(stream))
        check(status, "pvt_intersect_device_packed")

    def intersect(self, d_positions, d_directions, n, d_t0, d_hit, d_container, d_adjacent, stream=0):
        status = self._lib.pvt_intersect_device(self._handle, C.c_void_p(d_positions), C.c_void_p(d_directions), int(n),
                                                C.c_void_p(d_t0), C.c_void_p(d_hit), C.c_void_p(d_container),
                                                C.c_void_p(d_adjacent), C.c_void_p(stream))
        check(status, "pvt_intersect_device")
