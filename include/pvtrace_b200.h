/*
 * pvtrace_b200.h -- C ABI of the B200-native photon tracer.
 *
 * This is the drop-in boundary for ONE hot path of danieljfarrell/pvtrace: the call
 *
 *     pvtrace.engine._kernel.trace_bundle(compiled, positions, directions, wavelengths,
 *                                         seed, maxsteps, max_events, emit_method,
 *                                         num_threads, record_every) -> dict
 *
 * (reference: pvtrace/engine/_kernel.pyx:903-1115, called only from pvtrace/engine/api.py:233-244).
 * Everything here is plain C: pointers, sizes, PODs.  No torch / CUDA types appear in a signature; a
 * CUDA stream is passed as an opaque `void*` (0 = default stream).
 *
 * Conventions
 *   - every entry point returns 0 on success, non-zero on failure; pvt_last_error() then returns a
 *     thread-local, NUL-terminated description (CUDA error string, bad argument, ...).
 *   - "host" entry points take host pointers and perform the H2D/D2H copies themselves;
 *     "device" entry points take device pointers and only enqueue work on the given stream.
 *   - all floating point is IEEE binary64, all tables are C-contiguous, lengths are in elements.
 *   - there is NO CPU fallback in this library: without a CUDA device every compute entry fails loudly.
 */
#ifndef PVTRACE_B200_H
#define PVTRACE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVT_VERSION 200 /* 0.2.0 */

/* Compile-time limits shared with the reference kernel (_kernel.pyx:65-68, compiler.py:23). */
#define PVT_MAX_NODES 128
#define PVT_MAX_RECORDERS 256

/* Event codes == pvtrace.light.event.Event (light/event.py:7-16). */
enum {
  PVT_EV_GENERATE = 0, PVT_EV_REFLECT = 1, PVT_EV_TRANSMIT = 2, PVT_EV_ABSORB = 3, PVT_EV_NONRADIATIVE = 4,
  PVT_EV_SCATTER = 5, PVT_EV_EMIT = 6, PVT_EV_EXIT = 7, PVT_EV_REACT = 8, PVT_EV_KILL = 9
};
/* Table tags == pvtrace/engine/compiler.py:25-50. */
enum { PVT_GEOM_BOX = 0, PVT_GEOM_SPHERE = 1, PVT_GEOM_CYLINDER = 2 };
enum { PVT_SURF_FRESNEL = 0, PVT_SURF_NULL = 1 };
enum { PVT_COMP_ABSORBER = 0, PVT_COMP_SCATTERER = 1, PVT_COMP_LUMINOPHORE = 2, PVT_COMP_REACTOR = 3 };
enum { PVT_PHASE_ISOTROPIC = 0, PVT_PHASE_HENYEY_GREENSTEIN = 1, PVT_PHASE_CONE = 2 };
enum { PVT_EMIT_KT = 0, PVT_EMIT_REDSHIFT = 1, PVT_EMIT_FULL = 2 };
/* Recorder selectors == pvtrace/engine/recorder.py:45-53. */
enum {
  PVT_REC_ENTERING = 0, PVT_REC_ESCAPING = 1, PVT_REC_REFLECTED = 2, PVT_REC_LOST = 3, PVT_REC_REACTED = 4,
  PVT_REC_KILLED = 5, PVT_REC_EXIT = 6
};
/* Random stream selection.  PHILOX is the product default (counter based, north_star); XOSHIRO reproduces
 * the reference's splitmix64-seeded xoshiro256+ (_kernel.pyx:75-113) so per-ray histories can be diffed
 * against the compiled reference kernel. */
enum { PVT_RNG_PHILOX = 0, PVT_RNG_XOSHIRO = 1 };

/* Facet-surface extension flags (lowering of pvtrace/device/lsc.py:22-86 delegates into data). */
enum {
  PVT_FACET_TRANSMIT_STRAIGHT = 1, /* transmitted ray keeps its direction (index-matched solar cell, lsc.py:49-62) */
  PVT_FACET_REFLECT_LAMBERTIAN = 2 /* reflected direction is Lambertian about the outward normal (lsc.py:79-86)   */
};

/*
 * Flat scene tables.  Field-for-field the attributes of pvtrace.engine.compiler.CompiledScene
 * (compiler.py:57-204) that _kernel.trace_bundle reads (_kernel.pyx:933-1017), plus the facet-surface
 * extension at the end (n_facets == 0 reproduces the reference exactly).
 */
typedef struct pvt_scene_t {
  int32_t n_nodes;       /* <= PVT_MAX_NODES */
  int32_t root_id;
  int32_t n_components;
  int32_t n_abs_knots;   /* len(abs_x) == len(abs_y) */
  int32_t n_ems_knots;   /* len(ems_x) == len(ems_cdf) */
  int32_t n_recorders;   /* <= PVT_MAX_RECORDERS */
  int32_t n_hists;
  int32_t total_bins;
  int32_t n_facets;
  int32_t n_refl_knots;  /* len(refl_x) == len(refl_y): wavelength-tabulated facet reflectivities */

  const int32_t* geom_type;        /* [n_nodes]      PVT_GEOM_*                                        */
  const double*  geom_params;      /* [n_nodes,4]    box: sx,sy,sz,- | sphere: r | cylinder: length,radius */
  const double*  local_to_world;   /* [n_nodes,4,4]  row major, rigid                                  */
  const double*  world_to_local;   /* [n_nodes,4,4]                                                    */
  const double*  refractive_index; /* [n_nodes]                                                        */
  const int32_t* surface_type;     /* [n_nodes]      PVT_SURF_*                                        */
  const int32_t* comp_start;       /* [n_nodes]                                                        */
  const int32_t* comp_count;       /* [n_nodes]                                                        */

  const int32_t* comp_type;        /* [n_components] PVT_COMP_*                                        */
  const double*  comp_qy;
  const double*  comp_tau_rad;
  const double*  comp_tau_nr;
  const int32_t* comp_phase_type;  /* PVT_PHASE_*                                                      */
  const double*  comp_phase_param;
  const int32_t* comp_abs_start;
  const int32_t* comp_abs_n;
  const int32_t* comp_ems_start;
  const int32_t* comp_ems_n;
  const double*  abs_x;            /* [n_abs_knots]                                                    */
  const double*  abs_y;
  const double*  ems_x;            /* [n_ems_knots]                                                    */
  const double*  ems_cdf;

  const int32_t* rec_node;         /* [n_recorders]                                                    */
  const int32_t* rec_event;        /* PVT_REC_*                                                        */
  const int32_t* rec_has_facet;
  const double*  rec_facet;        /* [max(n_recorders,1),3] world-frame outward normal               */
  const double*  rec_atol;
  const int32_t* rec_hist_start;
  const int32_t* rec_hist_n;
  const int32_t* hist_prop_a;      /* [n_hists] property ids, recorder.py:33-41                        */
  const int32_t* hist_prop_b;      /* -1 => 1-D                                                        */
  const int32_t* hist_na;
  const int32_t* hist_nb;
  const double*  hist_lo_a;
  const double*  hist_hi_a;
  const double*  hist_lo_b;
  const double*  hist_hi_b;
  const int32_t* hist_offset;

  /* facet-surface extension: per node a list of facets, matched on the LOCAL outward normal */
  const int32_t* facet_start;        /* [n_nodes] (may be NULL when n_facets == 0)                     */
  const int32_t* facet_count;        /* [n_nodes]                                                      */
  const double*  facet_normal;       /* [n_facets,3]                                                   */
  const double*  facet_atol;         /* [n_facets] per-component tolerance                             */
  const double*  facet_reflectivity; /* [n_facets] constant R in [0,1]; < 0 => Fresnel                 */
  const int32_t* facet_flags;        /* [n_facets] PVT_FACET_*                                         */
  /* coatings (examples/006 Coatings.ipynb): a facet may cover only PART of a face -- the open box
   * lo < p < hi of the node's local frame -- and its reflectivity may be a table over wavelength (linear
   * interpolation, clamped to the end values and to [0,1]), which then replaces facet_reflectivity.      */
  const double*  facet_region;       /* [n_facets,6] lo xyz, hi xyz (+-inf = unbounded); NULL => whole faces */
  const int32_t* facet_refl_start;   /* [n_facets] into refl_x / refl_y; NULL => no tables              */
  const int32_t* facet_refl_n;       /* [n_facets] knots, 0 => facet_reflectivity applies               */
  const double*  refl_x;             /* [n_refl_knots] nanometres, ascending                            */
  const double*  refl_y;             /* [n_refl_knots] reflectivity                                     */
} pvt_scene_t;

/* On-device emission of the built-in light delegates (pvtrace/light/light.py:48-157,
 * pvtrace/engine/emit.py:22-89).  Rays cycle through lights as `index % n_lights` (scene/scene.py:141-151). */
enum { PVT_LPOS_POINT = 0, PVT_LPOS_RECT = 1, PVT_LPOS_CIRCLE = 2, PVT_LPOS_CUBE = 3 };
enum { PVT_LDIR_Z = 0, PVT_LDIR_CONE = 1, PVT_LDIR_ISOTROPIC = 2, PVT_LDIR_LAMBERTIAN = 3, PVT_LDIR_HG = 4 };
enum { PVT_LWL_CONSTANT = 0, PVT_LWL_SPECTRUM = 1 };

typedef struct pvt_emit_t {
  int32_t n_lights;
  int32_t n_wl_knots;
  const double*  light_to_world; /* [n_lights,4,4]                                        */
  const int32_t* pos_kind;       /* [n_lights] PVT_LPOS_*                                 */
  const double*  pos_param;      /* [n_lights,3] half extents / radius                    */
  const int32_t* dir_kind;       /* [n_lights] PVT_LDIR_*                                 */
  const double*  dir_param;      /* [n_lights] theta_max | g                              */
  const int32_t* wl_kind;        /* [n_lights] PVT_LWL_*                                  */
  const double*  wl_param;       /* [n_lights] nanometres for CONSTANT                    */
  const int32_t* wl_start;       /* [n_lights] into wl_x / wl_cdf for SPECTRUM            */
  const int32_t* wl_n;
  const double*  wl_x;           /* [n_wl_knots]                                          */
  const double*  wl_cdf;
} pvt_emit_t;

typedef struct pvt_params_t {
  int64_t  n;            /* rays in this bundle                                                          */
  int64_t  first_index;  /* global index of ray 0: ray i draws from stream id (seed + first_index + i),   */
                         /* mirroring the reference's `seed + i` (_kernel.pyx:1090, api.py:252-262)      */
  uint64_t seed;
  int64_t  record_every; /* 0: no event log; k: rays with i % k == 0 get a full history                   */
  int32_t  maxsteps;
  int32_t  max_events;
  int32_t  emit_method;  /* PVT_EMIT_*                                                                    */
  int32_t  rng_mode;     /* PVT_RNG_*                                                                     */
  int32_t  device;       /* CUDA ordinal (host entry points only)                                         */
  int32_t  flags;        /* PVT_FLAG_* bits                                                               */
} pvt_params_t;

/* Kernel selection.  By default a bundle is traced by the shared-memory wavefront kernel whenever the scene
 * allows it (Philox stream, <= 64 recorders, tables + photon pool fit in shared memory) and by the
 * one-photon-per-lane register kernel otherwise; this flag forces the latter (tests cross-check the two). */
enum { PVT_FLAG_REGISTER_KERNEL = 1 };

/* Run statistics written by every trace (device counters, not estimates). */
enum {
  PVT_STAT_STEPS = 0,      /* trace-loop iterations == "photon steps" (SURVEY 8d unit of work)          */
  PVT_STAT_RAYS = 1,       /* rays retired                                                               */
  PVT_STAT_LAUNCHES = 2,   /* kernels launched by the call                                               */
  PVT_STAT_EVENTS = 3,     /* events generated (logged or not)                                           */
  PVT_STAT_H2D_BYTES = 4,  /* host entry points: bytes of ray data that crossed PCIe (constant columns do not) */
  PVT_NSTATS = 32
};

/* Outputs of one bundle: exactly the dict returned by the reference's trace_bundle (_kernel.pyx:1097-1115).
 * The caller allocates everything.  rows = ceil(n / record_every) * max_events (0 when record_every == 0);
 * log pointers may be NULL when rows == 0.  Tallies are OVERWRITTEN (not accumulated) by host entry points. */
typedef struct pvt_out_t {
  int32_t* counts;        /* [ceil(n/record_every)] events logged per recorded ray */
  int64_t* rec_distinct;  /* [n_recorders]                                         */
  int64_t* rec_crossings; /* [n_recorders]                                         */
  double*  rec_sums;      /* [n_recorders,4,2]                                     */
  int64_t* rec_bins;      /* [total_bins]                                          */
  uint8_t* kind;          /* [rows]                                                */
  int32_t* hit;           /* [rows] (-1 filled)                                    */
  int32_t* container;
  int32_t* adjacent;
  int32_t* component;
  int32_t* source;
  double*  position;      /* [rows,3]                                              */
  double*  direction;
  double*  normal;
  double*  wavelength;    /* [rows]                                                */
  double*  travelled;
  double*  duration;
  int64_t* stats;         /* [PVT_NSTATS] or NULL                                  */
} pvt_out_t;

/* ---------------------------------------------------------------- library / device queries ----------- */
int         pvt_version(void);
int         pvt_device_count(void);      /* 0 when no CUDA device/driver is usable */
const char* pvt_last_error(void);
/* sizeof(pvt_scene_t), sizeof(pvt_emit_t), sizeof(pvt_params_t), sizeof(pvt_out_t): lets a foreign-function
 * binding verify its struct mirrors against the compiled library */
void        pvt_struct_sizes(int32_t sizes[4]);

/* Measured fp64 FMA throughput of the device in TFLOP/s (2 flops per FMA): eight independent chains per thread on every
 * scheduler.  bench.py reports the tracer's fp64 rate against it (the bound that governs it is issue latency, not HBM). */
int         pvt_measure_fp64_peak(int device, double* tflops);

/* ---------------------------------------------------------------- drop-in trace (host buffers) -------- *
 * Replaces _kernel.trace_bundle (_kernel.pyx:903-1115).  `positions`/`directions` are [n,3], `wavelengths`
 * [n] host arrays and are not modified (the reference copies them, :1064-1065).  When `emit` is non-NULL the
 * three arrays may be NULL and initial rays are sampled on the device instead (replaces
 * pvtrace.engine.emit.emit_bundle, emit.py:92-134).  `elapsed_s` receives the device time of the whole call
 * including the H2D/D2H copies (CUDA events), matching EngineResult.elapsed (api.py:232-245).               */
int pvt_trace_bundle(const pvt_scene_t* scene, const pvt_emit_t* emit,
                     const double* positions, const double* directions, const double* wavelengths,
                     const pvt_params_t* params, pvt_out_t* out, double* elapsed_s);

/* The same bundle over several devices of this process (`n_devices, device_ids`): contiguous index
 * slices, one host thread per device, per-device cached contexts; tallies are summed and the log rows of every slice
 * land in place, so `out` is what one device would have produced (integer fields identical, sums to summation order).
 * `params->device` is ignored.  This is what `engine.simulate(..., workers=k)` maps to (reference: `workers` =
 * OpenMP threads, pvtrace/engine/api.py:197-246).                                                              */
int pvt_trace_bundle_devices(const pvt_scene_t* scene, const pvt_emit_t* emit,
                             const double* positions, const double* directions, const double* wavelengths,
                             const pvt_params_t* params, int32_t n_devices, const int32_t* device_ids,
                             pvt_out_t* out, double* elapsed_s);

/* ---------------------------------------------------------------- resident-scene API ------------------ *
 * A context owns the device copy of the tables, the tally accumulators and the event-log buffers on ONE
 * device, so bundles can be streamed (engine.simulate_stream, api.py:249-264) without re-uploading.        */
typedef struct pvt_context pvt_context_t;

int pvt_context_create(const pvt_scene_t* scene, const pvt_emit_t* emit, int device, pvt_context_t** ctx);
int pvt_context_destroy(pvt_context_t* ctx);
/* zero the device tally accumulators + stats (async on stream) */
int pvt_context_reset(pvt_context_t* ctx, void* stream);
/* Trace rays whose initial state is ALREADY in device memory (d_* are device pointers; all three may be NULL
 * when the context has an emitter).  Tallies accumulate into the context.  Event-log device buffers are
 * (re)allocated by the context when record_every > 0 and fetched by pvt_context_read.  Asynchronous.       */
int pvt_trace_device(pvt_context_t* ctx, const double* d_positions, const double* d_directions,
                     const double* d_wavelengths, const pvt_params_t* params, void* stream);
/* synchronise `stream` and copy tallies (+ the event log of the LAST trace) to host memory */
int pvt_context_read(pvt_context_t* ctx, pvt_out_t* out, void* stream);
/* device pointer + element count of the packed tally buffer (doubles; integers are exactly representable):
 * [distinct n_rec | crossings n_rec | sums n_rec*8 | bins total_bins].  This is the buffer a multi-GPU caller
 * all-reduces (one ncclAllReduce(sum, f64)) before pvt_context_read_packed.                                */
int pvt_context_pack_tallies(pvt_context_t* ctx, double** d_packed, int64_t* n_packed, void* stream);
int pvt_context_unpack_tallies(pvt_context_t* ctx, void* stream);
/* Sample n initial rays on the device into caller-provided device arrays ([n,3],[n,3],[n]). */
int pvt_emit_device(pvt_context_t* ctx, double* d_positions, double* d_directions, double* d_wavelengths,
                    int64_t n, int64_t first_index, uint64_t seed, void* stream);
/* Host convenience wrapper around pvt_emit_device (replaces emit.emit_bundle for built-in delegates). */
int pvt_emit_bundle(const pvt_emit_t* emit, double* positions, double* directions, double* wavelengths,
                    int64_t n, int64_t first_index, uint64_t seed, int device);

/* ---------------------------------------------------------------- intersect stage on its own ---------- *
 * Stage kernel "ray-primitive intersect over a photon SoA" (next_hit + find_container,
 * photon_tracer.py:26-109 == _kernel.pyx:666-714): for each ray the nearest forward hit distance t0 and the
 * (hit, container, adjacent) node ids (-1 when nothing is hit).  Host pointers.                             */
int pvt_intersect_bundle(const pvt_scene_t* scene, const double* positions, const double* directions,
                         int64_t n, double* t0, int32_t* hit, int32_t* container, int32_t* adjacent,
                         int device, double* elapsed_s);
/* device-pointer variants.  The packed form is the stage as SURVEY 8d counts it against the HBM roofline -- 60 B per
 * ray: 48 in, t0 (8) + ONE word of ids (4) out, ids = hit | container << 8 | adjacent << 16 with 0xff for "none" (node
 * indices are < PVT_MAX_NODES); the other writes the three int32 arrays of the host call (68 B per ray).          */
int pvt_intersect_device(pvt_context_t* ctx, const double* d_positions, const double* d_directions, int64_t n,
                         double* d_t0, int32_t* d_hit, int32_t* d_container, int32_t* d_adjacent, void* stream);
int pvt_intersect_device_packed(pvt_context_t* ctx, const double* d_positions, const double* d_directions, int64_t n,
                                double* d_t0, uint32_t* d_ids, void* stream);

/* ---------------------------------------------------------------- device math helpers (known-answer tests)
 * One thin kernel launch each; host pointers; `n` independent evaluations.  They expose the SAME device
 * functions the tracer uses, so the reference's unit tests can be run against device code.                 */
/* material/utils.py:8-22 == _kernel.pyx:406-419 */
int pvt_test_fresnel_reflectivity(int64_t n, const double* angle, const double* n1, const double* n2, double* out, int device);
/* material/utils.py:25-32 == _kernel.pyx:422-433; d,nrm,out are [n,3] */
int pvt_test_specular_reflect(int64_t n, const double* d, const double* nrm, double* out, int device);
/* material/utils.py:35-45 == _kernel.pyx:436-446 (normal is flipped along the ray first, surface.py:165-171) */
int pvt_test_fresnel_refract(int64_t n, const double* d, const double* nrm, const double* n1, const double* n2, double* out, int device);
/* _kernel.pyx:245-345: geom_type[n], params[n,4], o[n,3], d[n,3] -> nhit[n], ts[n,4] (local frame, t > EPS) */
int pvt_test_intersect(int64_t n, const int32_t* geom_type, const double* params, const double* o, const double* d,
                       int32_t* nhit, double* ts, int device);
/* _kernel.pyx:359-400: outward local normal at local point p */
int pvt_test_local_normal(int64_t n, const int32_t* geom_type, const double* params, const double* p, double* out, int device);
/* _kernel.pyx:219-238 (np.interp with edge clamping) over one table xs/ys[m] */
int pvt_test_interp(int64_t n, const double* x, int32_t m, const double* xs, const double* ys, double* out, int device);
/* the device's own arithmetic helpers (no reference counterpart: the reference calls libm's log and the C division,
   _kernel.pyx:746-760): op 0 log(a) for normal positive a, 1 a / b, 2 1 / a, 3 sqrt(a) for a in [0, 1],
   as the trace kernels compute them */
int pvt_test_math(int64_t n, int32_t op, const double* a, const double* b, double* out, int device);
/* draws[n_rays, n_draws] of the per-ray uniform stream (rng_mode, seed + first_index + i) */
int pvt_test_rng_uniform(int64_t n_rays, int32_t n_draws, uint64_t seed, int64_t first_index, int32_t rng_mode, double* out, int device);
/* _kernel.pyx:455-476 using draws from the ray's own stream: out[n,3] */
int pvt_test_sample_phase(int64_t n, int32_t phase_type, double phase_param, uint64_t seed, int32_t rng_mode, double* out, int device);

#ifdef __cplusplus
}
#endif
#endif /* PVTRACE_B200_H */
