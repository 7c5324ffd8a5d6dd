// pvt_photon.cuh -- the physics of one photon step, split into the stages the kernels schedule:
//
//   classify_step   next_hit + find_container + Beer-Lambert free path  -> EXIT | VOLUME | SURFACE | dead
//   volume_step     absorption: component pick, re-emission (phase function, spectral CDF) or loss
//   surface_step    normal, Fresnel (or facet) reflectivity, reflect / refract
//   exit_step       leaves through the root boundary
//
// Behaviour follows the reference's compiled tracer, pvtrace/engine/_kernel.pyx:603-897 (which replicates
// pvtrace/algorithm/photon_tracer.py:112-273); citations on each block point at the lines reproduced.  The
// code is written for the GPU, not transcribed: hits are reduced on the fly instead of stored and sorted,
// scene records are shared-memory broadcasts, angles are carried as cosines (acos only when a recorder asks),
// tallies are fire-and-forget reductions, and every random number has a fixed counter address (pvt_rng.cuh).
#pragma once
#include "pvt_math.cuh"
#include "pvt_rng.cuh"
#include "pvt_scene.cuh"

namespace pvt {

typedef unsigned long long u64;

// Event-log columns in device memory (pvt_out_t without the tallies)
struct LogColumns {
  int32_t* counts;
  uint8_t* kind;
  int32_t *hit, *container, *adjacent, *component, *source;
  double *position, *direction, *normal, *wavelength, *travelled, *duration;
};

// Where tallies go.  Per-recorder scalars (distinct, crossings, 8 moment sums) go to a slab PRIVATE TO THE CTA in
// global memory, written with reductions that do not return (RED): no CAS loops, no contention between CTAs;
// the kernel folds the slab into the context's accumulators when the CTA retires.  Histogram bins are spread
// addresses and go straight to the shared accumulator.
struct TallySink {
  u64* distinct;  // [R]   CTA slab
  u64* cross;     // [R]
  double* sums;   // [R,8]
  u64* bins;      // [total_bins] context accumulator
};

struct StepParams {
  int maxsteps, max_events, emit_method;
};

// Photon state carried between steps.  `seen` is the distinct-ray mask of up to 32 * kSeenWords recorders.
// An accumulator that lives in (shared) memory: read, assigned and added to in place.  The wavefront kernel keeps a
// photon's path length and time of flight in their pool columns this way instead of in four registers that are live
// through the whole interaction just to be stored back at its end.
struct MemAcc {
  double* at;
  __device__ __forceinline__ operator double() const { return *at; }
  __device__ __forceinline__ MemAcc& operator=(double v) { *at = v; return *this; }
  __device__ __forceinline__ MemAcc& operator+=(double v) { *at += v; return *this; }
};

template <int kSeenWords, class Acc = double>
struct PhotonT {
  V3 p, d;
  double wl;
  Acc travelled, duration;
  int32_t source, count, nlog;
  int32_t log_ray;     // ordinal of this ray among the recorded ones (index into counts), < 0: not sampled
  long long log_base;  // first log row of this ray == log_ray * max_events, < 0 when the ray is not sampled
  uint32_t seen[kSeenWords];
};

enum StepClass { kDead = 0, kExit = 1, kVolume = 2, kSurface = 3, kKill = 4 };  // kKill: step budget exhausted

// next_hit + find_container (photon_tracer.py:26-109 == _kernel.pyx:666-714) as a single pass over the nodes that
// keeps the two nearest roots overall and the nearest root among nodes hit exactly once.  The reduction is branch
// free (selects on validity flags); ties resolve to the lowest node index (strict '<'), like the reference's scans.
struct Nearest {
  double t0;
  int hit, container, adjacent, total;
};

struct TwoNearest {
  double t_first, t_second, t_single;
  int n_first, n_second, n_single;
  __device__ __forceinline__ TwoNearest()
      : t_first(PVT_INF), t_second(PVT_INF), t_single(PVT_INF), n_first(-1), n_second(-1), n_single(-1) {}
  // the reference's insertion: a root becomes the nearest if it beats it (or is the first one), else the second
  // nearest if it beats that (or is the second one); +inf sentinels make "is the first / second one" implicit
  __device__ __forceinline__ void add(double t, int node, bool ok) {
    const bool lt1 = ok && t < t_first;
    const bool lt2 = ok && !lt1 && t < t_second;
    t_second = lt1 ? t_first : (lt2 ? t : t_second);
    n_second = lt1 ? n_first : (lt2 ? node : n_second);
    t_first = lt1 ? t : t_first;
    n_first = lt1 ? node : n_first;
  }
  // Both roots of a box or a sphere at once: a <= b whenever both count, and a counts only if b does.  The same
  // outcome as add(a) followed by add(b) (ties keep the earlier entry, strict '<') with three compares.
  __device__ __forceinline__ void add_pair(double a, double b, int node, bool ok_a, bool ok_b) {
    const double x = ok_a ? a : b;  // the nearer valid root (valid iff ok_b), b being the farther one (valid iff ok_a)
    const bool x_first = ok_b && x < t_first;
    const bool y_first = ok_a && b < t_first;
    const bool x_second = ok_b && x < t_second;
    const double t_keep = x_second ? x : t_second, t_push = y_first ? b : t_first;
    const int n_keep = x_second ? node : n_second, n_push = y_first ? node : n_first;
    t_second = x_first ? t_push : t_keep;
    n_second = x_first ? n_push : n_keep;
    t_first = x_first ? x : t_first;
    n_first = x_first ? node : n_first;
    add_single(b, node, ok_b && !ok_a);  // one root: it is the exit
  }
  // a node with exactly one root: candidate container
  __device__ __forceinline__ void add_single(double t, int node, bool ok) {
    const bool lt = ok && t < t_single;
    t_single = lt ? t : t_single;
    n_single = lt ? node : n_single;
  }
  // hit = nearest root's node; container = the nearest node hit exactly once, else the hit node; adjacent = the other of
  // the two nearest (:684-714)
  __device__ __forceinline__ Nearest result() const {
    Nearest r;
    r.total = n_first < 0 ? 0 : (n_second < 0 ? 1 : 2);  // 0, 1, "2 or more"
    r.t0 = t_first;
    r.hit = n_first;
    if (r.total <= 1) {
      r.container = n_first;
      r.adjacent = -1;
    } else {
      r.container = n_single >= 0 ? n_single : n_first;
      r.adjacent = r.container == n_first ? n_second : n_first;
    }
    return r;
  }
};

__device__ __forceinline__ Nearest nearest_surface(const SceneView& sv, const V3& p, const V3& d) {
  const int n_nodes = sv.hdr().n_nodes;
  TwoNearest best;
  V3 inv_world;           // 1 / d, shared by every axis-aligned box (their local direction IS d)
  bool have_inv = false;
  for (int node = 0; node < n_nodes; ++node) {
    const double* rec = sv.node(node);
    const int gtype = sv.node_int(node, NI_GEOM);
    V3 o, dl;
    if (sv.node_int(node, NI_ALIGNED)) {
      // rotation part of w2l is the identity: o = p + translation exactly as the full product would give
      o = V3{p.x + rec[kNodeW2L + 3], p.y + rec[kNodeW2L + 7], p.z + rec[kNodeW2L + 11]};
      dl = d;
    } else {
      o = map_point(rec + kNodeW2L, p);
      dl = map_vector(rec + kNodeW2L, d);
    }
    if (gtype == 0) {
      V3 inv;
      if (sv.node_int(node, NI_ALIGNED)) {
        if (!have_inv) { inv_world = slab_reciprocal(d); have_inv = true; }
        inv = inv_world;
      } else {
        inv = slab_reciprocal(dl);
      }
      double t_in, t_out;
      bool ok_in, ok_out;
      box_roots(rec[kNodeHalf], rec[kNodeHalf + 1], rec[kNodeHalf + 2], o, dl, inv, t_in, t_out, ok_in, ok_out);
      best.add_pair(t_in, t_out, node, ok_in, ok_out);
    } else if (gtype == 1) {
      double t1, t2;
      bool ok1, ok2;
      sphere_roots(rec[kNodeParams], o, dl, t1, t2, ok1, ok2);
      best.add_pair(t1, t2, node, ok1, ok2);
    } else {
      const Roots r = cylinder_roots(rec[kNodeParams], rec[kNodeParams + 1], o, dl);
      int count = 0;
      double only = PVT_INF;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        best.add(r.t[k], node, r.ok[k]);
        count += r.ok[k] ? 1 : 0;
        only = r.ok[k] ? r.t[k] : only;
      }
      best.add_single(only, node, count == 1);
    }
  }
  return best.result();
}

// The same reduction for scenes made of axis-aligned boxes only (every LSC scene: world box, slab, coatings), chosen
// per scene by the kernels' kBoxes instantiation: no primitive switch, no rotation, one reciprocal direction for all.
static __device__ __noinline__ Nearest nearest_surface_boxes_any(const SceneView sv, V3 p, V3 d) {  // rays parallel to a slab: rare
  const int n_nodes = sv.hdr().n_nodes;
  TwoNearest best;
  const V3 inv = slab_reciprocal(d);
  const double* rec = sv.node(0);
  for (int node = 0; node < n_nodes; ++node, rec += kNodeWords) {
    const V3 o = V3{p.x + rec[kNodeW2L + 3], p.y + rec[kNodeW2L + 7], p.z + rec[kNodeW2L + 11]};
    double t_in, t_out;
    bool ok_in, ok_out;
    box_roots(rec[kNodeHalf], rec[kNodeHalf + 1], rec[kNodeHalf + 2], o, d, inv, t_in, t_out, ok_in, ok_out);
    best.add_pair(t_in, t_out, node, ok_in, ok_out);
  }
  return best.result();
}
// kLean: the instruction-lean forms of the parallel test and the reciprocal (pvt_math.cuh), for the intersect stage
#ifndef PVT_TRACE_LEAN_RCP
#define PVT_TRACE_LEAN_RCP 1  // the Newton reciprocal in the trace kernels too: -1.3 to -2 % on the LSC configs (one lease)
#endif
template <bool kLean = false, bool kLeanRcp = kLean || PVT_TRACE_LEAN_RCP != 0>
__device__ __forceinline__ Nearest nearest_surface_boxes(const SceneView& sv, const V3& p, const V3& d) {
  if (kLean ? slab_parallel_lean(d) : slab_parallel(d)) return nearest_surface_boxes_any(sv, p, d);  // (the direction is the same in every node's frame)
  const int n_nodes = sv.hdr().n_nodes;
  const V3 inv = kLeanRcp ? slab_reciprocal_lean(d) : slab_reciprocal(d);
  const double* rec = sv.node(0);
  TwoNearest best;
  {  // node 0 straight into the empty reduction: what add_pair would select against +inf sentinels
    const V3 o = V3{p.x + rec[kNodeW2L + 3], p.y + rec[kNodeW2L + 7], p.z + rec[kNodeW2L + 11]};
    double t_in, t_out;
    bool ok_in, ok_out;
    box_roots_oblique(rec[kNodeHalf], rec[kNodeHalf + 1], rec[kNodeHalf + 2], o, d, inv, t_in, t_out, ok_in, ok_out);
    best.t_first = ok_in ? t_in : (ok_out ? t_out : PVT_INF);
    best.n_first = ok_out ? 0 : -1;  // ok_in implies ok_out
    best.t_second = ok_in ? t_out : PVT_INF;
    best.n_second = ok_in ? 0 : -1;
    best.t_single = (ok_out && !ok_in) ? t_out : PVT_INF;
    best.n_single = (ok_out && !ok_in) ? 0 : -1;
  }
  rec += kNodeWords;
  for (int node = 1; node < n_nodes; ++node, rec += kNodeWords) {
    const V3 o = V3{p.x + rec[kNodeW2L + 3], p.y + rec[kNodeW2L + 7], p.z + rec[kNodeW2L + 11]};
    double t_in, t_out;
    bool ok_in, ok_out;
    box_roots_oblique(rec[kNodeHalf], rec[kNodeHalf + 1], rec[kNodeHalf + 2], o, d, inv, t_in, t_out, ok_in, ok_out);
    best.add_pair(t_in, t_out, node, ok_in, ok_out);
  }
  return best.result();
}

// interpolate component c's absorption table at wavelength x
__device__ __forceinline__ double absorption_at(const SceneView& sv, int c, double x) {
  const Header& h = sv.hdr();
  const int start = sv.comp_int(c, CI_ABS_START), n = sv.comp_int(c, CI_ABS_N);
  return interp_hinted(x, sv.w + h.off_abs_x + start, sv.w + h.off_abs_y + start, n, sv.comp(c)[kCompAbsInvDx]);
}

// ---- event log (_kernel.pyx:562-597).  Out of line and by value: only sampled rays ever get here. ----------

static __device__ __noinline__ void log_event(const LogColumns& L, long long row, int kind, int hit, int container, int adjacent,
                                       int component, int source, V3 p, V3 d, bool has_normal, V3 normal, double wl,
                                       double travelled, double duration) {
  L.kind[row] = (uint8_t)kind;
  L.hit[row] = hit; L.container[row] = container; L.adjacent[row] = adjacent;
  L.component[row] = component; L.source[row] = source;
  L.position[3 * row] = p.x; L.position[3 * row + 1] = p.y; L.position[3 * row + 2] = p.z;
  L.direction[3 * row] = d.x; L.direction[3 * row + 1] = d.y; L.direction[3 * row + 2] = d.z;
  if (has_normal) { L.normal[3 * row] = normal.x; L.normal[3 * row + 1] = normal.y; L.normal[3 * row + 2] = normal.z; }
  L.wavelength[row] = wl; L.travelled[row] = travelled; L.duration[row] = duration;
}
// log one event of photon `ph` if the ray is sampled and has budget left; `kLog` is a compile-time flag of the
// enclosing function: kernels instantiated for record_every == 0 carry no logging code at all
#define PVT_LOG_N(ph, kind, hit, cont, adj, comp, has_n, nrm)                                                          \
  do {                                                                                                                 \
    if (kLog && (ph).log_base >= 0 && (ph).nlog < sp.max_events) {                                                     \
      log_event(L, (ph).log_base + (ph).nlog, kind, hit, cont, adj, comp, (ph).source, (ph).p, (ph).d, has_n, nrm,     \
                (ph).wl, (ph).travelled, (ph).duration);                                                               \
      ++(ph).nlog;                                                                                                     \
    }                                                                                                                  \
  } while (0)
#define PVT_LOG(ph, kind, hit, cont, adj, comp) PVT_LOG_N(ph, kind, hit, cont, adj, comp, false, (V3{0.0, 0.0, 0.0}))

// ---- tallies (_kernel.pyx:501-556; semantics restated by engine/tally.py:26-47,86-156) -------------------

// reductions into GLOBAL memory that return nothing (SASS RED): no round trip, no CAS loop
__device__ __forceinline__ void red_add(u64* p, u64 v) {
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ double ray_property(int prop, double wl, double angle, double duration, double travelled,
                                               const V3& lp) {
  double v = lp.z;
  v = prop == 0 ? wl : v;
  v = prop == 1 ? angle : v;
  v = prop == 2 ? duration : v;
  v = prop == 3 ? travelled : v;
  v = prop == 4 ? lp.x : v;
  v = prop == 5 ? lp.y : v;
  return v;
}

template <int SW>
struct SeenMask {
  uint32_t w[SW];
};

// `cosine` is the cosine of the incidence angle (1 for volume events => angle 0); acos is taken only when a
// recorder matches for the first time.  The candidates are the recorders attached to (node, selector) -- usually
// none to six -- looked up in the blob's index; lanes search independently and update together.
template <int SW>
__device__ __forceinline__ SeenMask<SW> tally_event(const SceneView& sv, const TallySink& T, SeenMask<SW> seen, int sel,
                                                    int node, int face, bool has_normal, const V3& wnormal,
                                                    const V3& lp, double cosine, double wl, double duration,
                                                    double travelled) {
  // Box faces have their matching recorders resolved on the host (the facet test against the face's world normal
  // does not depend on the event); everything else searches the recorders of (node, selector).
  int k, k_end = -1;
  if (face >= 0) sv.face_range(node, sel, face, k, k_end);
  const bool resolved = k_end >= 0;
  if (!resolved) sv.rec_range(node, sel, k, k_end);
  k_end += k;
  double angle = -1.0;
  for (;;) {
    int r = -1;
    if (resolved) {
      if (k < k_end) r = sv.face_candidate(k);
    } else {
      for (; k < k_end; ++k) {
        const int cand = sv.rec_candidate(k);
        if (!sv.rec_int(cand, RI_HAS_FACET)) { r = cand; break; }
        if (!has_normal) continue;
        const double* q = sv.rec(cand);
        const double tol = q[kRecAtol];
        if (fabs(q[0] - wnormal.x) <= tol && fabs(q[1] - wnormal.y) <= tol && fabs(q[2] - wnormal.z) <= tol) { r = cand; break; }
      }
    }
    if (r < 0) break;
    ++k;
    red_add(&T.cross[r], 1ull);
    bool was_seen;
    if (SW == 2) {  // 64 recorders: two scalar words, no indexing
      const uint32_t bit = 1u << (r & 31);
      const bool high = r >= 32;
      was_seen = ((high ? seen.w[1] : seen.w[0]) & bit) != 0;
      seen.w[0] |= high ? 0u : bit;
      seen.w[SW - 1] |= high ? bit : 0u;
    } else {
      const uint32_t bit = 1u << (r & 31);
      was_seen = false;
#pragma unroll
      for (int w = 0; w < SW; ++w)
        if (w == (r >> 5)) { was_seen = (seen.w[w] & bit) != 0; seen.w[w] |= bit; }
    }
    if (was_seen) continue;
    if (angle < 0.0) angle = cosine >= 1.0 ? 0.0 : acos(cosine);
    red_add(&T.distinct[r], 1ull);
    double* m = T.sums + 8 * r;
    red_add(m + 0, wl);         red_add(m + 1, wl * wl);
    red_add(m + 2, angle);      red_add(m + 3, angle * angle);
    red_add(m + 4, duration);   red_add(m + 5, duration * duration);
    red_add(m + 6, travelled);  red_add(m + 7, travelled * travelled);
    const int h0 = sv.rec_int(r, RI_HIST_START), h1 = h0 + sv.rec_int(r, RI_HIST_N);
    for (int h = h0; h < h1; ++h) {
      const double* g = sv.hist(h);
      const int na = sv.hist_int(h, HI_NA), nb = sv.hist_int(h, HI_NB), pb = sv.hist_int(h, HI_PROP_B);
      const double va = ray_property(sv.hist_int(h, HI_PROP_A), wl, angle, duration, travelled, lp);
      const int ia = (int)((va - g[kHistLoA]) / (g[kHistHiA] - g[kHistLoA]) * na);
      if (ia < 0 || ia >= na) continue;
      int bin = ia;
      if (pb >= 0) {
        const double vb = ray_property(pb, wl, angle, duration, travelled, lp);
        const int ib = (int)((vb - g[kHistLoB]) / (g[kHistHiB] - g[kHistLoB]) * nb);
        if (ib < 0 || ib >= nb) continue;
        bin = ia * nb + ib;
      }
      red_add(&T.bins[sv.hist_int(h, HI_OFFSET) + bin], 1ull);
    }
  }
  return seen;
}

// A stage does not tally itself: it fills in a request and the kernel makes the single (inlined) call once the
// photon has been written back, when almost nothing is live in registers.
struct TallyReq {
  int sel = -1;  // PVT_REC_*, < 0: nothing to tally
  int node;
  int face = -1;  // box face of a surface event (pre-resolved recorder list), -1: search with the facet test
  bool has_normal;
  V3 normal, lp;
  double cosine;
};

template <class P>
__device__ __forceinline__ void tally(const SceneView& sv, const TallySink& T, P& ph, const TallyReq& tr) {
  constexpr int SW = (int)(sizeof(ph.seen) / 4);
  SeenMask<SW> seen;
#pragma unroll
  for (int k = 0; k < SW; ++k) seen.w[k] = ph.seen[k];
  seen = tally_event<SW>(sv, T, seen, tr.sel, tr.node, tr.face, tr.has_normal, tr.normal, tr.lp, tr.cosine, ph.wl,
                         ph.duration, ph.travelled);
#pragma unroll
  for (int k = 0; k < SW; ++k) ph.seen[k] = seen.w[k];
}

// facet-surface extension: first facet of `node` whose LOCAL normal equals nl within its tolerance and whose region
// (an open box of the local frame; the whole space unless the facet is a partial coating) holds the local point
__device__ __forceinline__ int find_facet(const SceneView& sv, int node, const V3& nl, const V3& lp) {
  if (sv.hdr().n_facets <= 0) return -1;
  const int f0 = sv.node_int(node, NI_FACET_START), f1 = f0 + sv.node_int(node, NI_FACET_COUNT);
  for (int f = f0; f < f1; ++f) {
    const double* q = sv.facet(f);
    const double tol = q[kFacetAtol];
    if (fabs(q[0] - nl.x) <= tol && fabs(q[1] - nl.y) <= tol && fabs(q[2] - nl.z) <= tol &&
        lp.x > q[kFacetRegion] && lp.y > q[kFacetRegion + 1] && lp.z > q[kFacetRegion + 2] &&
        lp.x < q[kFacetRegion + 3] && lp.y < q[kFacetRegion + 4] && lp.z < q[kFacetRegion + 5])
      return f;
  }
  return -1;
}

// `slowness` = n / c of the container (seconds per cm), precomputed per node
template <class P>
__device__ __forceinline__ void advance(P& ph, double t, double slowness) {
  ph.p = axpy(ph.p, ph.d, t);
  ph.travelled += t;
  ph.duration += t * slowness;
}

// Per-lane run statistics (device counters of pvt_out_t.stats)
struct LaneStats {
  uint32_t steps = 0, events = 0, rays = 0;
  __device__ __forceinline__ void step() { ++steps; }
  __device__ __forceinline__ void event() { ++events; }
  __device__ __forceinline__ void ray() { ++rays; }
};
// The same counters as three words of the thread's own in SHARED memory, bumped by reductions that return nothing
// (RED.shared: one instruction, no register, nothing to wait for).  Counters that live for the whole kernel and change
// two or three times per photon step are the first thing a register allocator under pressure spills: as registers they
// were local-memory round trips in the wavefront kernel.  T = stride between the three arrays.
template <int T>
struct SmemStats {
  uint32_t at;  // shared-memory address of this thread's step counter
  __device__ __forceinline__ static void bump(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }
  __device__ __forceinline__ void step() { bump(at); }
  __device__ __forceinline__ void event() { bump(at + 4u * T); }
  __device__ __forceinline__ void ray() { bump(at + 8u * T); }
};

template <bool kLog, class P, class St>
__device__ __forceinline__ void begin_photon(P& ph, const LogColumns& L, const StepParams& sp, St& st) {
  ph.travelled = 0.0; ph.duration = 0.0;
  ph.source = -1; ph.count = 0; ph.nlog = 0;
#pragma unroll
  for (int w = 0; w < (int)(sizeof(ph.seen) / 4); ++w) ph.seen[w] = 0u;
  st.ray(); st.event();
  PVT_LOG(ph, PVT_EV_GENERATE, -1, -1, -1, -1);
}

// What classify_step hands to the second stage
struct StepPlan {
  double t;    // distance to the event: free path (VOLUME) or surface distance (SURFACE, EXIT)
  double u;    // the surface uniform of this step (address kBlockPath half 1); philox streams fetch it here
  double alpha;
  int hit, container, adjacent;
};

// First half of the reference's loop body (_kernel.pyx:654-760): budget check, intersect, kill check, free path.
// Touches nothing but position, direction, wavelength and the step counter.
// `u_slot`: where an addressed stream's surface uniform goes as soon as it is drawn -- the pool column of the wavefront
// kernel -- instead of staying live across the intersection just to be stored afterwards (plan.u is then not set).
template <bool kLog, bool kBoxes = false, class Rng, class P, class St>
__device__ __forceinline__ StepClass classify_step(const SceneView& sv, const LogColumns& L, const StepParams& sp, P& ph,
                                                   Rng& rng, St& st, StepPlan& plan, double* u_slot = nullptr) {
  const Header& H = sv.hdr();
  plan.u = 1.0; plan.alpha = 0.0;
  ++ph.count;
  rng.begin_step((uint32_t)ph.count);
  // event budget of sampled rays: keep room for the KILL record (:658-663)
  if (kLog && ph.log_base >= 0 && ph.nlog >= sp.max_events - 1) {
    st.event();
    PVT_LOG(ph, PVT_EV_KILL, -1, -1, -1, -1);
    return kDead;
  }
  st.step();
  // Addressed streams draw the step's two uniforms (free path, surface test) BEFORE the intersection: the ten Philox
  // rounds and the logarithm are chains that depend on nothing else, and issued here they overlap the slab arithmetic
  // instead of following it (the kernel is bound by dependent-issue latency, not by issue slots: -3 % on config 2).
  double ud_early = 0.0, u_early = 1.0;
  if (Rng::kAddressed) rng.pair(kBlockPath, ud_early, u_early);
  if (Rng::kAddressed && u_slot != nullptr) { *u_slot = u_early; u_early = 1.0; }
  const Nearest nh = kBoxes ? nearest_surface_boxes<true>(sv, ph.p, ph.d) : nearest_surface(sv, ph.p, ph.d);
  if (nh.total == 0) return kDead;  // :681-682
  plan.hit = nh.hit; plan.container = nh.container; plan.adjacent = nh.adjacent;
  plan.t = nh.t0;
  if (ph.count > sp.maxsteps) return kKill;  // :716-723, completed by kill_step
  if (nh.hit == H.root_id) return kExit;

  // Beer-Lambert free path in the container (material.py:17-47 == :746-760)
  // (Computing the coefficient and the free path AHEAD of the intersection -- possible where one node alone has
  // components, as in every LSC: they depend on the wavelength and the draw only -- shortens the dependent chain and
  // was measured 8 % SLOWER: two more doubles live across the slab tests, spilled.  Registers bound this kernel.)
  const int c0 = sv.node_int(nh.container, NI_COMP_START), cn = sv.node_int(nh.container, NI_COMP_COUNT);
  double alpha = 0.0;
  for (int k = 0; k < cn; ++k) alpha += absorption_at(sv, c0 + k, ph.wl);
  plan.alpha = alpha;
  if (alpha > kAlphaZero) {
    double ud;
    if (Rng::kAddressed) {
      ud = ud_early; plan.u = u_early;
    } else {
      ud = rng.one(kBlockPath, 0);
    }
    const double depth = PVT_DIV(-PVT_LOG1M(ud), alpha);
    if (depth < nh.t0) {
      plan.t = depth;
      return kVolume;
    }
  } else if (Rng::kAddressed) {
    plan.u = u_early;
  }
  return kSurface;
}

// Step budget exhausted (:716-723): KILL record and `killed` tally on the container, no movement.
template <bool kLog, class P, class St>
__device__ __forceinline__ void kill_step(const SceneView& sv, const LogColumns& L, const StepParams& sp, P& ph,
                                          St& st, const StepPlan& plan, TallyReq& tr) {
  st.event();
  PVT_LOG(ph, PVT_EV_KILL, -1, plan.container, -1, -1);
  if (sv.hdr().n_recorders > 0 && sv.has_recorders(plan.container, PVT_REC_KILLED)) {
    tr.sel = PVT_REC_KILLED; tr.node = plan.container; tr.has_normal = false; tr.normal = V3{0.0, 0.0, 0.0};
    tr.lp = map_point(sv.node(plan.container) + kNodeW2L, ph.p); tr.cosine = 1.0;
  }
}

// Leaves the scene through the root boundary (:728-744).
// kBoxes (scenes of axis-aligned boxes only, as in nearest_surface_boxes): the local point is a translation away, the
// normal is +-e_ax in both frames and every product with it is a component pick -- the values the general expressions
// give (multiplications by exact zeros and ones), without forming them.
template <bool kLog, bool kBoxes = false, class P, class St>
__device__ __forceinline__ void exit_step(const SceneView& sv, const LogColumns& L, const StepParams& sp, P& ph,
                                          St& st, const StepPlan& plan, TallyReq& tr) {
  const int hit = plan.hit;
  advance(ph, plan.t, sv.node(plan.container)[kNodeSlowness]);
  st.event();
  PVT_LOG(ph, PVT_EV_EXIT, hit, plan.container, plan.adjacent, -1);
  if (sv.hdr().n_recorders > 0 && sv.has_recorders(hit, PVT_REC_EXIT)) {
    const double* rec = sv.node(hit);
    V3 lp, nw;
    int face;
    double c;
    if (kBoxes) {
      lp = V3{ph.p.x + rec[kNodeW2L + 3], ph.p.y + rec[kNodeW2L + 7], ph.p.z + rec[kNodeW2L + 11]};
      face = box_face(rec + kNodeParams, lp);
      nw = axis_vector(face >> 1, (face & 1) ? 1.0 : -1.0);
      c = fabs(component(ph.d, face >> 1));
    } else {
      lp = map_point(rec + kNodeW2L, ph.p);
      const V3 nl = outward_normal(sv.node_int(hit, NI_GEOM), rec + kNodeParams, lp, face);
      nw = map_vector(rec + kNodeL2W, nl);
      c = fabs(dot(nw, ph.d));
    }
    if (c > 1.0) c = 1.0;
    tr.sel = PVT_REC_EXIT; tr.node = hit; tr.face = face; tr.has_normal = true; tr.normal = nw; tr.lp = lp; tr.cosine = c;
  }
}

// Absorbed in the volume (:762-832).  Returns true while the photon lives (re-emitted or scattered).
template <bool kLog, class Rng, class P, class St>
__device__ __forceinline__ bool volume_step(const SceneView& sv, const LogColumns& L, const StepParams& sp, P& ph,
                                            Rng& rng, St& st, const StepPlan& plan, TallyReq& tr) {
  const Header& H = sv.hdr();
  const int container = plan.container;
  advance(ph, plan.t, sv.node(container)[kNodeSlowness]);
  const int c0 = sv.node_int(container, NI_COMP_START), cn = sv.node_int(container, NI_COMP_COUNT);
  double u_target, u_yield;
  // addressed streams: all three blocks of the step up front, as independent chains (see classify_step)
  double g1_early = 0.0, g2_early = 0.0, u_gamma_early = 0.0, u_delay_early = 0.0;
  if (Rng::kAddressed) {
    rng.pair(kBlockAbsorb, u_target, u_yield);
    rng.pair(kBlockPhase, g1_early, g2_early);
    rng.pair(kBlockEmit, u_gamma_early, u_delay_early);
  } else {
    u_target = rng.one(kBlockAbsorb, 0);
  }
  const double target = u_target * plan.alpha;
  double running = 0.0;
  int comp = c0;
  for (int k = 0; k < cn; ++k) {
    running += absorption_at(sv, c0 + k, ph.wl);
    if (target <= running) { comp = c0 + k; break; }
  }
  st.event();
  PVT_LOG(ph, PVT_EV_ABSORB, -1, container, -1, comp);
  const double* cr = sv.comp(comp);
  const int ctype = sv.comp_int(comp, CI_TYPE);
  bool radiative = false;
  if (ctype == PVT_COMP_SCATTERER || ctype == PVT_COMP_LUMINOPHORE) {
    if (!Rng::kAddressed) u_yield = rng.one(kBlockAbsorb, 1);
    radiative = u_yield < cr[kCompQy];
  }
  if (radiative) {
    double g1, g2;
    if (Rng::kAddressed) { g1 = g1_early; g2 = g2_early; }
    else rng.pair(kBlockPhase, g1, g2);
    ph.d = phase_direction(sv.comp_int(comp, CI_PHASE), cr[kCompPhaseParam], g1, g2);
    ph.source = comp;
    st.event();
    if (ctype == PVT_COMP_LUMINOPHORE) {  // component.py:381-440 == :795-812
      const int es = sv.comp_int(comp, CI_EMS_START), en = sv.comp_int(comp, CI_EMS_N);
      const double* ex = sv.w + H.off_ems_x + es;
      const double* ec = sv.w + H.off_ems_cdf + es;
      double p1 = 0.0;
      if (sp.emit_method != PVT_EMIT_FULL) {
        double nm = ph.wl;
        if (sp.emit_method == PVT_EMIT_KT) nm = PVT_DIV(1240.0, PVT_DIV(1240.0, nm) + 1.5 * kBoltzmannEv * 300.0);
        p1 = interp_hinted(nm, ex, ec, en, cr[kCompEmsInvDx]);
      }
      double u_gamma, u_delay;
      if (Rng::kAddressed) { u_gamma = u_gamma_early; u_delay = u_delay_early; }
      else u_gamma = rng.one(kBlockEmit, 0);
      const double gamma = p1 + (1.0 - p1) * u_gamma;
      ph.wl = sv.comp_int(comp, CI_HAS_GUIDE) ? interp_guided(gamma, ec, ex, en, sv.ems_guide(comp), kGuideBuckets)
                                              : interp(gamma, ec, ex, en);
      if (cr[kCompTauRad] > 0.0) {
        if (!Rng::kAddressed) u_delay = rng.one(kBlockEmit, 1);
        ph.duration += -PVT_LOG1M(u_delay) * cr[kCompTauRad];
      }
      PVT_LOG(ph, PVT_EV_EMIT, -1, container, -1, comp);
    } else {
      PVT_LOG(ph, PVT_EV_SCATTER, -1, container, -1, comp);
    }
    return true;
  }
  if (cr[kCompTauNr] > 0.0) ph.duration += -PVT_LOG1M(rng.one_rare(kBlockEmit, 1)) * cr[kCompTauNr];
  st.event();
  int sel;
  if (ctype == PVT_COMP_REACTOR) {
    PVT_LOG(ph, PVT_EV_REACT, -1, container, -1, comp);
    sel = PVT_REC_REACTED;
  } else {
    PVT_LOG(ph, PVT_EV_NONRADIATIVE, -1, container, -1, comp);
    sel = PVT_REC_LOST;
  }
  if (H.n_recorders > 0 && sv.has_recorders(container, sel)) {
    const V3 lp = map_point(sv.node(container) + kNodeW2L, ph.p);
    tr.sel = sel; tr.node = container; tr.has_normal = false; tr.normal = V3{0.0, 0.0, 0.0}; tr.lp = lp; tr.cosine = 1.0;
  }
  return false;
}

// Reaches a surface that is not the root boundary (:834-895).  Returns true while the photon lives.
template <bool kLog, bool kBoxes = false, class Rng, class P, class St>
__device__ __forceinline__ bool surface_step(const SceneView& sv, const LogColumns& L, const StepParams& sp, P& ph,
                                             Rng& rng, St& st, const StepPlan& plan, TallyReq& tr) {
  const int hit = plan.hit, container = plan.container, adjacent = plan.adjacent;
  const double n1 = sv.node(container)[kNodeIndex];
  advance(ph, plan.t, sv.node(container)[kNodeSlowness]);
  st.event();
  if (adjacent < 0) {
    PVT_LOG(ph, PVT_EV_KILL, hit, container, -1, -1);
    return false;
  }
  const double* hrec = sv.node(hit);
  V3 lp, nl, nw, nf;
  int face, ax = 0;
  double c, sg = 1.0, sf = 1.0, d_ax = 0.0;  // kBoxes: nl = nw = sg e_ax, nf = sf e_ax, d_ax = the direction's ax component
  if (kBoxes) {
    lp = V3{ph.p.x + hrec[kNodeW2L + 3], ph.p.y + hrec[kNodeW2L + 7], ph.p.z + hrec[kNodeW2L + 11]};
    face = box_face(hrec + kNodeParams, lp);
    ax = face >> 1; sg = (face & 1) ? 1.0 : -1.0;
    d_ax = component(ph.d, ax);
    sf = sg * d_ax < 0.0 ? -sg : sg;  // the normal turned along the ray
    c = sf * d_ax;
  } else {
    lp = map_point(hrec + kNodeW2L, ph.p);
    nl = outward_normal(sv.node_int(hit, NI_GEOM), hrec + kNodeParams, lp, face);
    nw = map_vector(hrec + kNodeL2W, nl);
    nf = nw;
    if (dot(nf, ph.d) < 0.0) nf = neg(nf);
    c = dot(nf, ph.d);  // cosine of the incidence angle, in [0, 1] up to rounding
  }
  const double c_raw = c;
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);

  const bool fresnel = sv.node_int(hit, NI_SURF) == PVT_SURF_FRESNEL;
  const double n2 = sv.node(adjacent)[kNodeIndex];
  double R = 0.0;
  bool straight = false, lambert = false, fixed_R = false;
  const int facet = sv.hdr().n_facets > 0 ? find_facet(sv, hit, kBoxes ? axis_vector(ax, sg) : nl, lp) : -1;
  if (facet >= 0) {
    const int flags = sv.facet_flags(facet);
    straight = (flags & PVT_FACET_TRANSMIT_STRAIGHT) != 0;
    lambert = (flags & PVT_FACET_REFLECT_LAMBERTIAN) != 0;
    double fr = sv.facet(facet)[kFacetRefl];
    const int rn = sv.facet_refl_n(facet);
    if (rn > 0) {  // coating with a reflectivity spectrum
      const int rs = sv.facet_refl_start(facet);
      fr = interp(ph.wl, sv.w + sv.hdr().off_refl_x + rs, sv.w + sv.hdr().off_refl_y + rs, rn);
      fr = fr < 0.0 ? 0.0 : (fr > 1.0 ? 1.0 : fr);
    }
    if (fr >= 0.0) { R = fr; fixed_R = true; }
    // a coating does not repeal Snell's law: where no refracted ray exists (total internal reflection) a facet that
    // transmits by refraction reflects, whatever reflectivity it states (the square root in snell() would be of a
    // negative number)
    if (fixed_R && fresnel && !straight && n2 < n1 && PVT_SQRT(fmax(1.0 - c * c, 0.0)) * PVT_DIV(n1, n2) > 1.0) R = 1.0;
  }
  if (!fixed_R && fresnel) R = fresnel_R_cos(c, n1, n2);

  double u = 1.0;
  if (R > 0.0) u = Rng::kAddressed ? plan.u : rng.one(kBlockPath, 1);  // surface.py:231-240: no draw when R == 0
  int sel;
  bool record = sv.hdr().n_recorders > 0;
  if (u < R) {
    if (lambert) {
      double p1, p2;
      rng.pair_rare(kBlockLambert, p1, p2);
      ph.d = lambert_about(kBoxes ? axis_vector(ax, -sf) : neg(nf), p1, p2);
    } else if (kBoxes) {
      ph.d = with_component(ph.d, ax, -d_ax);  // mirror(): d - 2 (d . n) n with n = +-e_ax
    } else {
      ph.d = mirror(ph.d, nw);
    }
    PVT_LOG_N(ph, PVT_EV_REFLECT, hit, container, adjacent, -1, true, (kBoxes ? axis_vector(ax, sg) : nw));
    sel = PVT_REC_REFLECTED;
    record = record && container != hit;
  } else {
    if (fresnel && !straight) {
      if (kBoxes) {  // snell(): n d + f nf with nf = sf e_ax and d . nf = c_raw >= 0
        const double n = PVT_DIV(n1, n2);
        const double f = PVT_SQRT(1.0 - n * n * (1.0 - c_raw * c_raw)) - n * c_raw;
        ph.d = with_component(V3{n * ph.d.x, n * ph.d.y, n * ph.d.z}, ax, n * d_ax + f * sf);
      } else {
        ph.d = snell(ph.d, nf, n1, n2);
      }
    }
    PVT_LOG_N(ph, PVT_EV_TRANSMIT, hit, container, adjacent, -1, true, (kBoxes ? axis_vector(ax, sg) : nw));
    sel = container == hit ? PVT_REC_ESCAPING : PVT_REC_ENTERING;
  }
  if (record && sv.has_recorders(hit, sel)) {
    tr.sel = sel; tr.node = hit; tr.face = face; tr.has_normal = true; tr.normal = kBoxes ? axis_vector(ax, sg) : nw; tr.lp = lp; tr.cosine = c;
  }
  return true;
}

// ---- on-device emission of the built-in light delegates (emit.py:22-134; scene.py:141-151) ---------------
// Uniform k of Philox stream kStreamEmit of photon `index` of run `run`: k=0 wavelength, k=1..3 position, k=4,5 direction.
struct EmittedRay {
  V3 pos, dir;
  double wl;
};
__device__ __forceinline__ EmittedRay emit_ray_value(const SceneView& sv, const RunSeed& run, long long index);

// out-of-line forms: by reference (refill fallback, register kernel, emit_kernel) and straight into the columns of a
// shared-memory ring of stride K (so the caller keeps nothing address-taken on its stack)
static __device__ __noinline__ void emit_ray(const SceneView sv, const RunSeed& run, long long index, V3& pos, V3& dir, double& wl) {
  const EmittedRay r = emit_ray_value(sv, run, index);
  pos = r.pos; dir = r.dir; wl = r.wl;
}
static __device__ __noinline__ void emit_ray_to_ring(const SceneView sv, const RunSeed& run, long long index, double* dst, int K) {
  const EmittedRay r = emit_ray_value(sv, run, index);
  dst[0] = r.pos.x; dst[K] = r.pos.y; dst[2 * K] = r.pos.z;
  dst[3 * K] = r.dir.x; dst[4 * K] = r.dir.y; dst[5 * K] = r.dir.z;
  dst[6 * K] = r.wl;
}

__device__ __forceinline__ EmittedRay emit_ray_value(const SceneView& sv, const RunSeed& run, long long index) {
  V3 pos, dir;
  double wl;
  const Header& H = sv.hdr();
  const u64 id = run.philox_base + (u64)index;
  const int l = (int)(index % H.n_lights);
  const double* q = sv.light(l);
  V3 lp = V3{0.0, 0.0, 0.0}, ld = V3{0.0, 0.0, 1.0};
  const int wl_kind = sv.light_int(l, LI_WL), pos_kind = sv.light_int(l, LI_POS);
  int kind = sv.light_int(l, LI_DIR);
  if (wl_kind == PVT_LWL_SPECTRUM || pos_kind != PVT_LPOS_POINT) {
    const U4 b0 = philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), 0u, kStreamEmit}, kPhiloxKey0, kPhiloxKey1);
    const U4 b1 = philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), 1u, kStreamEmit}, kPhiloxKey0, kPhiloxKey1);
    const double k0 = u53(b0.x, b0.y), k1 = u53(b0.z, b0.w), k2 = u53(b1.x, b1.y), k3 = u53(b1.z, b1.w);
    if (wl_kind == PVT_LWL_SPECTRUM) {
      const int s = sv.light_int(l, LI_WL_START), n = sv.light_int(l, LI_WL_N);
      wl = interp(k0, sv.w + H.off_wl_cdf + s, sv.w + H.off_wl_x + s, n);
    }
    const double px = q[kLightPos], py = q[kLightPos + 1], pz = q[kLightPos + 2];
    switch (pos_kind) {
      case PVT_LPOS_RECT:
        lp.x = -px + (px - -px) * k1; lp.y = -py + (py - -py) * k2;
        break;
      case PVT_LPOS_CIRCLE: {
        double s, c;
        sincospi(2.0 * k1, &s, &c);
        const double r = sqrt(k2) * px;
        lp.x = r * c; lp.y = r * s;
      } break;
      case PVT_LPOS_CUBE:
        lp.x = -px + (px - -px) * k1; lp.y = -py + (py - -py) * k2; lp.z = -pz + (pz - -pz) * k3;
        break;
      default: break;
    }
  }
  if (wl_kind != PVT_LWL_SPECTRUM) wl = q[kLightWl];
  const double prm = q[kLightDir];
  if (kind == PVT_LDIR_HG && fabs(prm) < 1e-12) kind = PVT_LDIR_ISOTROPIC;
  if (kind != PVT_LDIR_Z) {
    const U4 b2 = philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), 2u, kStreamEmit}, kPhiloxKey0, kPhiloxKey1);
    const double u0 = u53(b2.x, b2.y), u1 = u53(b2.z, b2.w);
    switch (kind) {
      case PVT_LDIR_CONE: {
        const double st = sqrt(u0) * q[kLightSinDir];  // sin(prm), taken once on the host
        ld = polar_sc(st, sqrt(fmax(1.0 - st * st, 0.0)), u1);
      } break;
      case PVT_LDIR_ISOTROPIC: {
        const double mu = 2.0 * u1 - 1.0;
        ld = polar_sc(sqrt(fmax(1.0 - mu * mu, 0.0)), mu, u0);
      } break;
      case PVT_LDIR_LAMBERTIAN: ld = polar_sc(sqrt(u0), sqrt(fmax(1.0 - u0, 0.0)), u1); break;
      case PVT_LDIR_HG: {
        const double s = 2.0 * u0 - 1.0;
        const double f = (1.0 - prm * prm) / (1.0 + prm * s);
        double mu = (1.0 + prm * prm - f * f) / (2.0 * prm);
        mu = mu > 1.0 ? 1.0 : (mu < -1.0 ? -1.0 : mu);
        ld = polar_sc(sqrt(1.0 - mu * mu), mu, u1);
      } break;
      default: break;
    }
  }
  pos = map_point(q + kLightL2W, lp);
  dir = map_vector(q + kLightL2W, ld);
  return EmittedRay{pos, dir, wl};
}

}  // namespace pvt
