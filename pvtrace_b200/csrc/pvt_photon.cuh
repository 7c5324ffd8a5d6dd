// pvt_photon.cuh -- one photon on one lane: state, nearest-surface search, tallies, event log, and the
// step function (one iteration of the trace loop = one "photon step", the unit of work of SURVEY 8d).
//
// Behaviour follows the reference's compiled tracer, pvtrace/engine/_kernel.pyx:603-897 (which replicates
// pvtrace/algorithm/photon_tracer.py:112-273); the citations on each block point at the lines reproduced.
// The code is written for the GPU, not transcribed: hits are reduced on the fly instead of being stored and
// sorted, scene records are read as shared-memory broadcasts, tallies go to a CTA-private shared slab and
// to global histogram bins with fire-and-forget atomics, and all random numbers come from a counter-based
// stream addressed by the photon's global index.
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

// Where tallies go: a shared-memory slab private to the CTA for the per-recorder scalars (every lane of the
// chip hammers the same few addresses otherwise) and global memory for histogram bins (spread addresses).
struct TallySink {
  u64* distinct;  // [R] shared
  u64* cross;     // [R] shared
  double* sums;   // [R,8] shared
  u64* bins;      // [total_bins] global
};

template <class Rng, int kSeenWords>
struct Photon {
  V3 p, d;
  double wl, travelled, duration;
  Rng rng;
  int32_t source, count, nlog;
  long long log_base;  // first log row of this ray, < 0 when the ray is not sampled
  uint32_t seen[kSeenWords];
  uint32_t nsteps, nevents;  // run statistics carried by the lane
};

struct Nearest {
  double t0;
  int hit, container, adjacent, total;
};

// next_hit + find_container (photon_tracer.py:26-109 == _kernel.pyx:666-714) as a single pass over the nodes:
// keeps the two nearest roots overall and the nearest root among nodes hit exactly once.  Ties resolve to the
// lowest node index (strict '<'), like the reference's scans.
__device__ __forceinline__ Nearest nearest_surface(const SceneView& sv, const V3& p, const V3& d) {
  const int n_nodes = sv.hdr().n_nodes;
  double t_first = PVT_INF, t_second = PVT_INF, t_single = PVT_INF;
  int n_first = -1, n_second = -1, n_single = -1, total = 0;
  for (int node = 0; node < n_nodes; ++node) {
    const double* rec = sv.node(node);
    const V3 o = map_point(rec + kNodeW2L, p);
    const V3 dl = map_vector(rec + kNodeW2L, d);
    double ts[4];
    const int k = roots(sv.node_int(node, NI_GEOM), rec + kNodeParams, o, dl, ts);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < k) {
        const double t = ts[j];
        if (total == 0 || t < t_first) {
          t_second = t_first; n_second = n_first;
          t_first = t; n_first = node;
        } else if (n_second < 0 || t < t_second) {
          t_second = t; n_second = node;
        }
        ++total;
      }
    }
    if (k == 1 && ts[0] < t_single) { t_single = ts[0]; n_single = node; }
  }
  Nearest r;
  r.total = total;
  r.t0 = t_first;
  r.hit = n_first;
  if (total <= 1) {
    r.container = n_first;
    r.adjacent = -1;
  } else {
    r.container = n_single >= 0 ? n_single : n_first;
    r.adjacent = r.container == n_first ? n_second : n_first;
  }
  return r;
}

// interpolate component c's absorption table at wavelength x
__device__ __forceinline__ double absorption_at(const SceneView& sv, int c, double x) {
  const Header& h = sv.hdr();
  const int start = sv.comp_int(c, CI_ABS_START), n = sv.comp_int(c, CI_ABS_N);
  return interp_hinted(x, sv.w + h.off_abs_x + start, sv.w + h.off_abs_y + start, n, sv.comp(c)[kCompAbsInvDx]);
}

// ---- event log (_kernel.pyx:562-597) ---------------------------------------------------------------------

template <class P>
__device__ __forceinline__ void log_event(const LogColumns& L, int max_events, P& ph, int kind, int hit, int container,
                                          int adjacent, int component, const V3* normal) {
  ++ph.nevents;
  if (ph.log_base < 0 || ph.nlog >= max_events) return;
  const long long row = ph.log_base + ph.nlog;
  L.kind[row] = (uint8_t)kind;
  L.hit[row] = hit; L.container[row] = container; L.adjacent[row] = adjacent;
  L.component[row] = component; L.source[row] = ph.source;
  L.position[3 * row] = ph.p.x; L.position[3 * row + 1] = ph.p.y; L.position[3 * row + 2] = ph.p.z;
  L.direction[3 * row] = ph.d.x; L.direction[3 * row + 1] = ph.d.y; L.direction[3 * row + 2] = ph.d.z;
  if (normal) { L.normal[3 * row] = normal->x; L.normal[3 * row + 1] = normal->y; L.normal[3 * row + 2] = normal->z; }
  L.wavelength[row] = ph.wl; L.travelled[row] = ph.travelled; L.duration[row] = ph.duration;
  ++ph.nlog;
}

// ---- tallies (_kernel.pyx:501-556; semantics restated by engine/tally.py:26-47,86-156) -------------------

__device__ __forceinline__ double ray_property(int prop, double wl, double angle, double duration, double travelled,
                                               const V3& lp) {
  switch (prop) {
    case 0: return wl;
    case 1: return angle;
    case 2: return duration;
    case 3: return travelled;
    case 4: return lp.x;
    case 5: return lp.y;
    default: return lp.z;
  }
}

template <class P>
__device__ __noinline__ void tally(const SceneView& sv, const TallySink& T, P& ph, int sel, int node,
                                   const V3* wnormal, const V3& lp, double angle) {
  const int R = sv.hdr().n_recorders;
  for (int r = 0; r < R; ++r) {
    if (sv.rec_int(r, RI_NODE) != node || sv.rec_int(r, RI_EVENT) != sel) continue;
    const double* q = sv.rec(r);
    if (sv.rec_int(r, RI_HAS_FACET)) {
      if (!wnormal) continue;
      const double tol = q[kRecAtol];
      if (fabs(q[0] - wnormal->x) > tol || fabs(q[1] - wnormal->y) > tol || fabs(q[2] - wnormal->z) > tol) continue;
    }
    atomicAdd(&T.cross[r], 1ull);
    const uint32_t bit = 1u << (r & 31);
    if (ph.seen[r >> 5] & bit) continue;
    ph.seen[r >> 5] |= bit;
    atomicAdd(&T.distinct[r], 1ull);
    double* m = T.sums + 8 * r;
    atomicAdd(m + 0, ph.wl);         atomicAdd(m + 1, ph.wl * ph.wl);
    atomicAdd(m + 2, angle);         atomicAdd(m + 3, angle * angle);
    atomicAdd(m + 4, ph.duration);   atomicAdd(m + 5, ph.duration * ph.duration);
    atomicAdd(m + 6, ph.travelled);  atomicAdd(m + 7, ph.travelled * ph.travelled);
    const int h0 = sv.rec_int(r, RI_HIST_START), h1 = h0 + sv.rec_int(r, RI_HIST_N);
    for (int h = h0; h < h1; ++h) {
      const double* g = sv.hist(h);
      const int na = sv.hist_int(h, HI_NA), nb = sv.hist_int(h, HI_NB), pb = sv.hist_int(h, HI_PROP_B);
      const double va = ray_property(sv.hist_int(h, HI_PROP_A), ph.wl, angle, ph.duration, ph.travelled, lp);
      const int ia = (int)((va - g[kHistLoA]) / (g[kHistHiA] - g[kHistLoA]) * na);
      if (ia < 0 || ia >= na) continue;
      int bin = ia;
      if (pb >= 0) {
        const double vb = ray_property(pb, ph.wl, angle, ph.duration, ph.travelled, lp);
        const int ib = (int)((vb - g[kHistLoB]) / (g[kHistHiB] - g[kHistLoB]) * nb);
        if (ib < 0 || ib >= nb) continue;
        bin = ia * nb + ib;
      }
      atomicAdd(&T.bins[sv.hist_int(h, HI_OFFSET) + bin], 1ull);
    }
  }
}

// facet-surface extension: first facet of `node` whose LOCAL normal equals nl within its tolerance
__device__ __forceinline__ int find_facet(const SceneView& sv, int node, const V3& nl) {
  if (sv.hdr().n_facets <= 0) return -1;
  const int f0 = sv.node_int(node, NI_FACET_START), f1 = f0 + sv.node_int(node, NI_FACET_COUNT);
  for (int f = f0; f < f1; ++f) {
    const double* q = sv.facet(f);
    const double tol = q[kFacetAtol];
    if (fabs(q[0] - nl.x) <= tol && fabs(q[1] - nl.y) <= tol && fabs(q[2] - nl.z) <= tol) return f;
  }
  return -1;
}

struct StepParams {
  int maxsteps, max_events, emit_method;
};

template <class P>
__device__ __forceinline__ void advance(P& ph, double t, double n_container) {
  ph.p = axpy(ph.p, ph.d, t);
  ph.travelled += t;
  ph.duration += t * n_container / kLightSpeed;
}

template <class Rng, int SW>
__device__ __forceinline__ void begin_photon(Photon<Rng, SW>& ph, const LogColumns& L, const StepParams& sp) {
  ph.travelled = 0.0; ph.duration = 0.0;
  ph.source = -1; ph.count = 0; ph.nlog = 0;
#pragma unroll
  for (int w = 0; w < SW; ++w) ph.seen[w] = 0u;
  log_event(L, sp.max_events, ph, PVT_EV_GENERATE, -1, -1, -1, -1, nullptr);
}

// One iteration of the reference's `while True` (_kernel.pyx:654-895).  Returns true while the photon lives.
template <class Rng, int SW>
__device__ __forceinline__ bool step_photon(const SceneView& sv, const TallySink& T, const LogColumns& L,
                                            const StepParams& sp, Photon<Rng, SW>& ph) {
  typedef Photon<Rng, SW> P;
  const Header& H = sv.hdr();
  const bool have_rec = H.n_recorders > 0;
  ++ph.count;
  // event budget of sampled rays: keep room for the KILL record (:658-663)
  if (ph.log_base >= 0 && ph.nlog >= sp.max_events - 1) {
    log_event(L, sp.max_events, ph, PVT_EV_KILL, -1, -1, -1, -1, nullptr);
    return false;
  }
  ++ph.nsteps;
  const Nearest nh = nearest_surface(sv, ph.p, ph.d);
  if (nh.total == 0) return false;  // :681-682
  const int hit = nh.hit, container = nh.container, adjacent = nh.adjacent;
  const double t0 = nh.t0;

  if (ph.count > sp.maxsteps) {  // :716-723
    log_event(L, sp.max_events, ph, PVT_EV_KILL, -1, container, -1, -1, nullptr);
    if (have_rec) {
      const V3 lp = map_point(sv.node(container) + kNodeW2L, ph.p);
      tally<P>(sv, T, ph, PVT_REC_KILLED, container, nullptr, lp, 0.0);
    }
    return false;
  }

  const double n_container = sv.node(container)[kNodeIndex];

  if (hit == H.root_id) {  // leaves the scene, :728-744
    advance(ph, t0, n_container);
    log_event(L, sp.max_events, ph, PVT_EV_EXIT, hit, container, adjacent, -1, nullptr);
    if (have_rec) {
      const double* rec = sv.node(hit);
      const V3 lp = map_point(rec + kNodeW2L, ph.p);
      const V3 nl = outward_normal(sv.node_int(hit, NI_GEOM), rec + kNodeParams, lp);
      const V3 nw = map_vector(rec + kNodeL2W, nl);
      double c = fabs(dot(nw, ph.d));
      if (c > 1.0) c = 1.0;
      tally<P>(sv, T, ph, PVT_REC_EXIT, hit, &nw, lp, acos(c));
    }
    return false;
  }

  // Beer-Lambert free path in the container (material.py:17-47 == :746-760)
  const int c0 = sv.node_int(container, NI_COMP_START), cn = sv.node_int(container, NI_COMP_COUNT);
  double alpha = 0.0;
  for (int k = 0; k < cn; ++k) alpha += absorption_at(sv, c0 + k, ph.wl);
  double depth = PVT_INF;
  if (alpha > kAlphaZero) depth = -log(1.0 - ph.rng.next()) / alpha;

  if (depth < t0) {  // absorbed in the volume, :762-832
    advance(ph, depth, n_container);
    const double target = ph.rng.next() * alpha;
    double running = 0.0;
    int comp = c0;
    for (int k = 0; k < cn; ++k) {
      running += absorption_at(sv, c0 + k, ph.wl);
      if (target <= running) { comp = c0 + k; break; }
    }
    log_event(L, sp.max_events, ph, PVT_EV_ABSORB, -1, container, -1, comp, nullptr);
    const double* cr = sv.comp(comp);
    const int ctype = sv.comp_int(comp, CI_TYPE);
    if ((ctype == PVT_COMP_SCATTERER || ctype == PVT_COMP_LUMINOPHORE) && ph.rng.next() < cr[kCompQy]) {
      ph.d = phase_direction(sv.comp_int(comp, CI_PHASE), cr[kCompPhaseParam], ph.rng);
      ph.source = comp;
      if (ctype == PVT_COMP_LUMINOPHORE) {  // component.py:381-440 == :795-812
        const int es = sv.comp_int(comp, CI_EMS_START), en = sv.comp_int(comp, CI_EMS_N);
        const double* ex = sv.w + H.off_ems_x + es;
        const double* ec = sv.w + H.off_ems_cdf + es;
        double p1 = 0.0;
        if (sp.emit_method != PVT_EMIT_FULL) {
          double nm = ph.wl;
          if (sp.emit_method == PVT_EMIT_KT) nm = 1240.0 / (1240.0 / nm + 1.5 * kBoltzmannEv * 300.0);
          p1 = interp_hinted(nm, ex, ec, en, cr[kCompEmsInvDx]);
        }
        const double gamma = p1 + (1.0 - p1) * ph.rng.next();
        ph.wl = interp(gamma, ec, ex, en);
        if (cr[kCompTauRad] > 0.0) ph.duration += -log(1.0 - ph.rng.next()) * cr[kCompTauRad];
        log_event(L, sp.max_events, ph, PVT_EV_EMIT, -1, container, -1, comp, nullptr);
      } else {
        log_event(L, sp.max_events, ph, PVT_EV_SCATTER, -1, container, -1, comp, nullptr);
      }
      return true;
    }
    if (cr[kCompTauNr] > 0.0) ph.duration += -log(1.0 - ph.rng.next()) * cr[kCompTauNr];
    int sel;
    if (ctype == PVT_COMP_REACTOR) {
      log_event(L, sp.max_events, ph, PVT_EV_REACT, -1, container, -1, comp, nullptr);
      sel = PVT_REC_REACTED;
    } else {
      log_event(L, sp.max_events, ph, PVT_EV_NONRADIATIVE, -1, container, -1, comp, nullptr);
      sel = PVT_REC_LOST;
    }
    if (have_rec) {
      const V3 lp = map_point(sv.node(container) + kNodeW2L, ph.p);
      tally<P>(sv, T, ph, sel, container, nullptr, lp, 0.0);
    }
    return false;
  }

  // reaches the surface, :834-895
  advance(ph, t0, n_container);
  if (adjacent < 0) {
    log_event(L, sp.max_events, ph, PVT_EV_KILL, hit, container, -1, -1, nullptr);
    return false;
  }
  const double* hrec = sv.node(hit);
  const V3 lp = map_point(hrec + kNodeW2L, ph.p);
  const V3 nl = outward_normal(sv.node_int(hit, NI_GEOM), hrec + kNodeParams, lp);
  const V3 nw = map_vector(hrec + kNodeL2W, nl);
  V3 nf = nw;
  if (dot(nf, ph.d) < 0.0) nf = neg(nf);
  double c = dot(nf, ph.d);
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  const double angle = acos(c);

  const bool fresnel = sv.node_int(hit, NI_SURF) == PVT_SURF_FRESNEL;
  const double n1 = n_container, n2 = sv.node(adjacent)[kNodeIndex];
  double R = 0.0;
  bool straight = false, lambert = false, fixed_R = false;
  const int facet = find_facet(sv, hit, nl);
  if (facet >= 0) {
    const int flags = sv.facet_flags(facet);
    straight = (flags & PVT_FACET_TRANSMIT_STRAIGHT) != 0;
    lambert = (flags & PVT_FACET_REFLECT_LAMBERTIAN) != 0;
    const double fr = sv.facet(facet)[kFacetRefl];
    if (fr >= 0.0) { R = fr; fixed_R = true; }
  }
  if (!fixed_R && fresnel) R = fresnel_R(angle, n1, n2);

  double u = 1.0;
  if (R > 0.0) u = ph.rng.next();  // surface.py:231-240: no draw when R == 0
  if (u < R) {
    ph.d = lambert ? lambert_about(neg(nf), ph.rng) : mirror(ph.d, nw);
    log_event(L, sp.max_events, ph, PVT_EV_REFLECT, hit, container, adjacent, -1, &nw);
    if (have_rec && container != hit) tally<P>(sv, T, ph, PVT_REC_REFLECTED, hit, &nw, lp, angle);
  } else {
    if (fresnel && !straight) ph.d = snell(ph.d, nf, n1, n2);
    log_event(L, sp.max_events, ph, PVT_EV_TRANSMIT, hit, container, adjacent, -1, &nw);
    if (have_rec) tally<P>(sv, T, ph, container == hit ? PVT_REC_ESCAPING : PVT_REC_ENTERING, hit, &nw, lp, angle);
  }
  return true;
}

// ---- on-device emission of the built-in light delegates (emit.py:22-134; scene.py:141-151) ---------------
// Draw k of Philox stream kStreamEmit of ray id: k=0 wavelength, k=1..3 position, k=4,5 direction.
__device__ __forceinline__ void emit_ray(const SceneView& sv, u64 id, long long index, V3& pos, V3& dir, double& wl) {
  const Header& H = sv.hdr();
  const int l = (int)(index % H.n_lights);
  const double* q = sv.light(l);
  V3 lp = V3{0.0, 0.0, 0.0}, ld = V3{0.0, 0.0, 1.0};
  const U4 b0 = philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), 0u, kStreamEmit}, kPhiloxKey0, kPhiloxKey1);
  const U4 b1 = philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), 1u, kStreamEmit}, kPhiloxKey0, kPhiloxKey1);
  const U4 b2 = philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), 2u, kStreamEmit}, kPhiloxKey0, kPhiloxKey1);
  const double k0 = u53(b0.x, b0.y), k1 = u53(b0.z, b0.w), k2 = u53(b1.x, b1.y), k3 = u53(b1.z, b1.w);
  const double u0 = u53(b2.x, b2.y), u1 = u53(b2.z, b2.w);
  if (sv.light_int(l, LI_WL) == PVT_LWL_SPECTRUM) {
    const int s = sv.light_int(l, LI_WL_START), n = sv.light_int(l, LI_WL_N);
    wl = interp(k0, sv.w + H.off_wl_cdf + s, sv.w + H.off_wl_x + s, n);
  } else {
    wl = q[kLightWl];
  }
  const double px = q[kLightPos], py = q[kLightPos + 1], pz = q[kLightPos + 2];
  switch (sv.light_int(l, LI_POS)) {
    case PVT_LPOS_RECT:
      lp.x = -px + (px - -px) * k1; lp.y = -py + (py - -py) * k2;
      break;
    case PVT_LPOS_CIRCLE: {
      double s, c;
      sincos(kTwoPi * k1, &s, &c);
      const double r = sqrt(k2) * px;
      lp.x = r * c; lp.y = r * s;
    } break;
    case PVT_LPOS_CUBE:
      lp.x = -px + (px - -px) * k1; lp.y = -py + (py - -py) * k2; lp.z = -pz + (pz - -pz) * k3;
      break;
    default: break;
  }
  const double prm = q[kLightDir];
  int kind = sv.light_int(l, LI_DIR);
  if (kind == PVT_LDIR_HG && fabs(prm) < 1e-12) kind = PVT_LDIR_ISOTROPIC;
  switch (kind) {
    case PVT_LDIR_CONE: ld = polar(asin(sqrt(u0) * sin(prm)), kTwoPi * u1); break;
    case PVT_LDIR_ISOTROPIC: ld = polar(acos(2.0 * u1 - 1.0), kTwoPi * u0); break;
    case PVT_LDIR_LAMBERTIAN: ld = polar(asin(sqrt(u0)), kTwoPi * u1); break;
    case PVT_LDIR_HG: {
      const double s = 2.0 * u0 - 1.0;
      const double f = (1.0 - prm * prm) / (1.0 + prm * s);
      const double mu = (1.0 + prm * prm - f * f) / (2.0 * prm);
      ld = polar(acos(mu), kTwoPi * u1);
    } break;
    default: break;
  }
  pos = map_point(q + kLightL2W, lp);
  dir = map_vector(q + kLightL2W, ld);
}

}  // namespace pvt
