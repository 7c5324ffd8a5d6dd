/*
 * pvt_oracle.c -- CPU ORACLE for the photon-tracing hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference algorithm, pvtrace/engine/_kernel.pyx (the compiled tracer, which
 * itself replicates pvtrace/algorithm/photon_tracer.py:112-273).  Each function cites the reference lines it
 * follows.  Nothing in the product library (pvtrace_b200/csrc) includes, links or calls this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PINNING: with rng_mode == PVT_RNG_XOSHIRO this oracle reproduces the compiled reference kernel
 * (oracle/_ref/_kernel*.so, built from /root/reference by `make -C oracle ref`) BIT FOR BIT -- event kinds,
 * node ids, positions, wavelengths, tallies -- see tests/test_oracle_pinning.py, and the committed fixtures
 * under tests/golden/ hold reference outputs for the same check where the reference tree is absent.
 * With rng_mode == PVT_RNG_PHILOX it consumes the same counter-based stream as the CUDA path, so per-ray
 * histories of the CUDA kernels can be diffed against it.
 *
 * Extensions beyond the reference (both are no-ops for reference-compatible input):
 *   - facet-surface table (scene->n_facets > 0): data-driven restatement of the Python delegates in
 *     pvtrace/device/lsc.py:22-86 (OptionalMirrorAndSolarCell, AirGapMirror/lambertian).
 *   - seeded emission of the built-in light delegates (pvtrace/engine/emit.py:22-134).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (no FMA contraction: keeps the arithmetic identical
 * to the reference build on x86-64).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/pvtrace_b200.h"

/* _kernel.pyx:29-34 */
static const double EPS_DIST = 2.220446049250313e-13;
static const double ALPHA_ZERO = 1e-8;
static const double C_CM_PER_S = 2.99792458e10;
#define KB_EV (1.380649e-23 / 1.60217662e-19)
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------
 * Random streams.
 *   XOSHIRO: splitmix64-seeded xoshiro256+, _kernel.pyx:82-113, state seeded with (seed + i) (:1090).
 *   PHILOX : Philox4x32-10 (Salmon et al., SC'11), fixed key, counter = (id lo, id hi, block, stream) with
 *            id = hash(seed) + first_index + i (splitmix64 finaliser: every seed owns its own window of the
 *            counter space); one block yields two 53-bit uniforms.  Shared definition with
 *            pvtrace_b200/csrc/pvt_rng.cuh.                                                                  */

#define PHILOX_KEY0 0x50565442u /* "PVTB" */
#define PHILOX_KEY1 0x32303042u /* "200B" */

static uint64_t hash_seed(uint64_t seed) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

static inline double u53(uint32_t lo, uint32_t hi) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

/* draw number `k` (0-based) of stream `stream` of photon `index` of the run seeded `seed` -- random access */
static double philox_uniform_at(uint64_t seed, uint64_t index, uint32_t stream, uint32_t k) {
  uint64_t id = hash_seed(seed) + index;
  uint32_t c[4] = {(uint32_t)id, (uint32_t)(id >> 32), k >> 1, stream};
  philox4x32_10(c, PHILOX_KEY0, PHILOX_KEY1);
  return (k & 1u) ? u53(c[2], c[3]) : u53(c[0], c[1]);
}

typedef struct {
  int mode;
  uint64_t s[4]; /* xoshiro state */
  uint64_t id;   /* philox: photon index within the run */
  uint64_t seed; /* philox: key */
  uint32_t k;    /* philox: sequential cursor of rng_next() (known-answer helpers only) */
  uint32_t step; /* philox: trace-loop iteration the addressed draws belong to */
} rng_t;

/* Draw addressing of the tracer in PHILOX mode (shared definition with pvtrace_b200/csrc/pvt_rng.cuh): every
 * random decision of a photon step has the fixed address (photon index, step, block, half) -> Philox counter
 * (index_lo, index_hi, step * 8 + block, stream 0), half selecting the first or second 53-bit uniform of the block.
 * In XOSHIRO mode addresses are ignored and draws are consumed in program order, which is the reference's. */
enum { BLOCK_PATH = 0, BLOCK_ABSORB = 1, BLOCK_PHASE = 2, BLOCK_EMIT = 3, BLOCK_LAMBERT = 4, BLOCKS_PER_STEP = 8 };

static uint64_t splitmix64(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static void rng_init(rng_t* r, int mode, uint64_t seed, uint64_t index) {
  r->mode = mode;
  r->id = index;
  r->seed = seed;
  r->k = 0;
  r->step = 0;
  uint64_t x = seed + index; /* the reference's per-ray seed, _kernel.pyx:1090 */
  for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&x);
}

static double rng_next(rng_t* r) {
  if (r->mode == PVT_RNG_XOSHIRO) {
    uint64_t* s = r->s;
    uint64_t result = s[0] + s[3];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = (s[3] << 45) | (s[3] >> 19);
    return (double)(result >> 11) * (1.0 / 9007199254740992.0);
  }
  return philox_uniform_at(r->seed, r->id, 0u, r->k++);
}

static double rng_draw(rng_t* r, uint32_t block, uint32_t half) {
  if (r->mode == PVT_RNG_XOSHIRO) return rng_next(r);
  return philox_uniform_at(r->seed, r->id, 0u, ((r->step * BLOCKS_PER_STEP + block) << 1) | half);
}

/* ------------------------------------------------------------------------------------------------
 * Small math: affine maps (_kernel.pyx:207-216), np.interp with clamping (:219-238).                */

static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static inline void xform_point(const double* m, const double* p, double* q) {
  for (int r = 0; r < 3; ++r) q[r] = m[4 * r] * p[0] + m[4 * r + 1] * p[1] + m[4 * r + 2] * p[2] + m[4 * r + 3];
}
static inline void xform_vector(const double* m, const double* v, double* q) {
  for (int r = 0; r < 3; ++r) q[r] = m[4 * r] * v[0] + m[4 * r + 1] * v[1] + m[4 * r + 2] * v[2];
}

/* last index i with xs[i] <= x, for xs[0] < x < xs[n-1]: what the reference's bisection (:228-235) converges to */
static double interp_clamped(double x, const double* xs, const double* ys, int n) {
  if (n == 1 || x <= xs[0]) return ys[0];
  if (x >= xs[n - 1]) return ys[n - 1];
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (xs[mid] <= x) lo = mid; else hi = mid;
  }
  if (xs[hi] == xs[lo]) return ys[lo];
  return ys[lo] + (ys[hi] - ys[lo]) * (x - xs[lo]) / (xs[hi] - xs[lo]);
}

/* ------------------------------------------------------------------------------------------------
 * Ray-primitive intersection in the local frame; returns roots with t > EPS (_kernel.pyx:245-345).      */

static int hit_box(const double* size, const double* o, const double* d, double* ts) {
  double tnear = -INFINITY, tfar = INFINITY;
  for (int ax = 0; ax < 3; ++ax) {
    double lo = -0.5 * size[ax], hi = 0.5 * size[ax];
    if (fabs(d[ax]) < 1e-300) {
      if (o[ax] < lo || o[ax] > hi) return 0;
      continue;
    }
    double inv = 1.0 / d[ax];
    double ta = (lo - o[ax]) * inv, tb = (hi - o[ax]) * inv;
    if (ta > tb) { double s = ta; ta = tb; tb = s; }
    if (ta > tnear) tnear = ta;
    if (tb < tfar) tfar = tb;
  }
  if (tfar < tnear) return 0;
  int n = 0;
  if (tnear > EPS_DIST) ts[n++] = tnear;
  if (tfar > EPS_DIST) ts[n++] = tfar;
  return n;
}

static int hit_sphere(double radius, const double* o, const double* d, double* ts) {
  double a = dot3(d, d);
  double b = 2.0 * dot3(d, o);
  double c = dot3(o, o) - radius * radius;
  double disc = b * b - 4.0 * a * c;
  if (disc < 0.0) return 0;
  double sq = sqrt(disc);
  int n = 0;
  double t = (-b - sq) / (2.0 * a);
  if (t > EPS_DIST) ts[n++] = t;
  t = (-b + sq) / (2.0 * a);
  if (t > EPS_DIST) ts[n++] = t;
  return n;
}

static int hit_cylinder(double length, double radius, const double* o, const double* d, double* ts) {
  double half = 0.5 * length;
  double cand[4];
  int nc = 0;
  double a = d[0] * d[0] + d[1] * d[1];
  if (a > 1e-300) { /* curved side, open interval in z */
    double b = 2.0 * (o[0] * d[0] + o[1] * d[1]);
    double c = o[0] * o[0] + o[1] * o[1] - radius * radius;
    double disc = b * b - 4.0 * a * c;
    if (disc >= 0.0) {
      double sq = sqrt(disc);
      double t = (-b - sq) / (2.0 * a);
      double z = o[2] + t * d[2];
      if (z > -half && z < half) cand[nc++] = t;
      t = (-b + sq) / (2.0 * a);
      z = o[2] + t * d[2];
      if (z > -half && z < half) cand[nc++] = t;
    }
  }
  if (fabs(d[2]) > 1e-300) { /* caps, closed discs */
    for (int s = 0; s < 2; ++s) {
      double zc = s == 0 ? -half : half;
      double t = (zc - o[2]) / d[2];
      double x = o[0] + t * d[0], y = o[1] + t * d[1];
      if (x * x + y * y <= radius * radius) cand[nc++] = t;
    }
  }
  int n = 0;
  for (int i = 0; i < nc; ++i)
    if (cand[i] > EPS_DIST) ts[n++] = cand[i];
  return n;
}

static int hit_primitive(int gtype, const double* prm, const double* o, const double* d, double* ts) {
  if (gtype == PVT_GEOM_BOX) return hit_box(prm, o, d, ts);
  if (gtype == PVT_GEOM_SPHERE) return hit_sphere(prm[0], o, d, ts);
  return hit_cylinder(prm[0], prm[1], o, d, ts);
}

/* Outward unit normal at local point p; never fails (_kernel.pyx:359-400). */
static void primitive_normal(int gtype, const double* prm, const double* p, double* n) {
  n[0] = n[1] = n[2] = 0.0;
  if (gtype == PVT_GEOM_BOX) {
    double best = INFINITY;
    int bax = 0, bsg = 1;
    for (int ax = 0; ax < 3; ++ax)
      for (int sg = -1; sg <= 1; sg += 2) {
        double dist = fabs(p[ax] - sg * 0.5 * prm[ax]);
        if (dist < best) { best = dist; bax = ax; bsg = sg; }
      }
    n[bax] = (double)bsg;
  } else if (gtype == PVT_GEOM_SPHERE) {
    double mag = sqrt(dot3(p, p));
    n[0] = p[0] / mag; n[1] = p[1] / mag; n[2] = p[2] / mag;
  } else {
    double half = 0.5 * prm[0];
    double tol = 1e-8 + 1e-5 * fabs(half); /* np.isclose defaults */
    if (fabs(p[2] + half) <= tol) n[2] = -1.0;
    else if (fabs(p[2] - half) <= tol) n[2] = 1.0;
    else {
      double r = sqrt(p[0] * p[0] + p[1] * p[1]);
      n[0] = p[0] / r; n[1] = p[1] / r;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Optics: material/utils.py:8-45 == _kernel.pyx:406-446; phase functions utils.py:104-186 == :455-476 */

static double fresnel_R(double angle, double n1, double n2) {
  if (n2 < n1 && angle > asin(n2 / n1)) return 1.0;
  double c = cos(angle), s = sin(angle);
  double q = n1 / n2 * s;
  double k = sqrt(1.0 - q * q);
  double rs = (n1 * c - n2 * k) / (n1 * c + n2 * k);
  double rp = (n1 * k - n2 * c) / (n1 * k + n2 * c);
  return 0.5 * (rs * rs + rp * rp);
}

static void mirror_dir(const double* d, const double* normal, double* out) {
  double n[3] = {normal[0], normal[1], normal[2]};
  if (dot3(n, d) < 0.0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  double dd = dot3(n, d);
  for (int i = 0; i < 3; ++i) out[i] = d[i] - 2.0 * dd * n[i];
}

/* `nf` already flipped to point along the ray */
static void snell_dir(const double* d, const double* nf, double n1, double n2, double* out) {
  double n = n1 / n2;
  double dd = dot3(d, nf);
  double c = sqrt(1.0 - n * n * (1.0 - dd * dd));
  double sign = dd < 0.0 ? -1.0 : 1.0;
  for (int i = 0; i < 3; ++i) out[i] = n * d[i] + sign * (c - sign * n * dd) * nf[i];
}

static void polar_dir(double theta, double phi, double* out) {
  out[0] = sin(theta) * cos(phi);
  out[1] = sin(theta) * sin(phi);
  out[2] = cos(theta);
}

/* g1, g2: the two uniforms of the phase-function draw, in the reference's order of consumption */
static void phase_dir(int ptype, double prm, double g1, double g2, double* out) {
  double theta, phi;
  if (ptype == PVT_PHASE_HENYEY_GREENSTEIN && fabs(prm) >= EPS_DIST) {
    double g = prm;
    double s = 2.0 * g1 - 1.0;
    double f = (1.0 - g * g) / (1.0 + g * s);
    double mu = 1.0 / (2.0 * g) * (1.0 + g * g - f * f);
    phi = 2.0 * M_PI * g2;
    theta = acos(mu);
  } else if (ptype == PVT_PHASE_CONE) {
    theta = asin(sqrt(g1) * sin(prm));
    phi = 2.0 * M_PI * g2;
  } else {
    phi = 2.0 * M_PI * g1;
    theta = acos(2.0 * g2 - 1.0);
  }
  polar_dir(theta, phi, out);
}

/* ------------------------------------------------------------------------------------------------
 * Tallies (_kernel.pyx:482-556; semantics independently stated by engine/tally.py:26-47,86-156).       */

typedef struct {
  int64_t* distinct; /* [R] */
  int64_t* cross;    /* [R] */
  double* sums;      /* [R,8] */
  int64_t* bins;     /* [total_bins] */
} acc_t;

static double ray_property(int prop, double wl, double angle, double duration, double travelled, const double* lp) {
  switch (prop) {
    case 0: return wl;
    case 1: return angle;
    case 2: return duration;
    case 3: return travelled;
    case 4: return lp[0];
    case 5: return lp[1];
    default: return lp[2];
  }
}

static void tally_event(const pvt_scene_t* S, acc_t* A, int sel, int node, unsigned char* seen,
                        const double* wnormal, const double* lp, double angle, double wl, double travelled,
                        double duration) {
  for (int r = 0; r < S->n_recorders; ++r) {
    if (S->rec_node[r] != node || S->rec_event[r] != sel) continue;
    if (S->rec_has_facet[r]) {
      if (!wnormal) continue;
      double tol = S->rec_atol[r];
      const double* f = S->rec_facet + 3 * r;
      if (fabs(f[0] - wnormal[0]) > tol || fabs(f[1] - wnormal[1]) > tol || fabs(f[2] - wnormal[2]) > tol) continue;
    }
    A->cross[r] += 1;
    if (seen[r]) continue;
    seen[r] = 1;
    A->distinct[r] += 1;
    double* m = A->sums + 8 * r;
    m[0] += wl;        m[1] += wl * wl;
    m[2] += angle;     m[3] += angle * angle;
    m[4] += duration;  m[5] += duration * duration;
    m[6] += travelled; m[7] += travelled * travelled;
    int h0 = S->rec_hist_start[r], h1 = h0 + S->rec_hist_n[r];
    for (int h = h0; h < h1; ++h) {
      double va = ray_property(S->hist_prop_a[h], wl, angle, duration, travelled, lp);
      int ia = (int)((va - S->hist_lo_a[h]) / (S->hist_hi_a[h] - S->hist_lo_a[h]) * S->hist_na[h]);
      if (ia < 0 || ia >= S->hist_na[h]) continue;
      if (S->hist_prop_b[h] < 0) {
        A->bins[S->hist_offset[h] + ia] += 1;
      } else {
        double vb = ray_property(S->hist_prop_b[h], wl, angle, duration, travelled, lp);
        int ib = (int)((vb - S->hist_lo_b[h]) / (S->hist_hi_b[h] - S->hist_lo_b[h]) * S->hist_nb[h]);
        if (ib < 0 || ib >= S->hist_nb[h]) continue;
        A->bins[S->hist_offset[h] + ia * S->hist_nb[h] + ib] += 1;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Event log (_kernel.pyx:562-597)                                                                    */

typedef struct {
  pvt_out_t* out;
  int64_t base; /* first row of this ray, < 0 when not sampled */
  int max_events;
  int n;        /* events written */
} log_t;

static void log_event(log_t* L, int kind, int hit, int container, int adjacent, int component, int source,
                      const double* pos, const double* dir, const double* normal, double wl, double travelled,
                      double duration) {
  if (L->base < 0 || L->n >= L->max_events) return;
  int64_t row = L->base + L->n;
  pvt_out_t* o = L->out;
  o->kind[row] = (uint8_t)kind;
  o->hit[row] = hit; o->container[row] = container; o->adjacent[row] = adjacent;
  o->component[row] = component; o->source[row] = source;
  for (int i = 0; i < 3; ++i) {
    o->position[3 * row + i] = pos[i];
    o->direction[3 * row + i] = dir[i];
    o->normal[3 * row + i] = normal ? normal[i] : 0.0;
  }
  o->wavelength[row] = wl; o->travelled[row] = travelled; o->duration[row] = duration;
  L->n += 1;
}

/* ------------------------------------------------------------------------------------------------
 * next_hit + find_container (photon_tracer.py:26-109 == _kernel.pyx:666-714), streaming over nodes.
 * Ties follow the reference: strict '<' so the earliest (lowest node index, then root order) wins.     */

typedef struct { double t0; int hit, container, adjacent, nhits; } nearest_t;

static void nearest_surface(const pvt_scene_t* S, const double* pos, const double* dir, nearest_t* R) {
  double t_first = INFINITY, t_second = INFINITY, t_single = INFINITY;
  int n_first = -1, n_second = -1, n_single = -1, total = 0;
  for (int node = 0; node < S->n_nodes; ++node) {
    double o[3], d[3], ts[4];
    xform_point(S->world_to_local + 16 * node, pos, o);
    xform_vector(S->world_to_local + 16 * node, dir, d);
    int k = hit_primitive(S->geom_type[node], S->geom_params + 4 * node, o, d, ts);
    double tmin = INFINITY;
    for (int j = 0; j < k; ++j) {
      double t = ts[j];
      if (t < tmin) tmin = t;
      if (total == 0 || t < t_first) {
        t_second = t_first; n_second = n_first;
        t_first = t; n_first = node;
      } else if (n_second < 0 || t < t_second) {
        t_second = t; n_second = node;
      }
      ++total;
    }
    if (k == 1 && tmin < t_single) { t_single = tmin; n_single = node; }
  }
  R->nhits = total;
  R->t0 = t_first;
  R->hit = n_first;
  if (total == 0) { R->container = R->adjacent = -1; return; }
  if (total == 1) { R->container = n_first; R->adjacent = -1; return; }
  int container = n_single >= 0 ? n_single : n_first;
  R->container = container;
  R->adjacent = container == n_first ? n_second : n_first;
}

/* ------------------------------------------------------------------------------------------------
 * One photon, start to finish (_kernel.pyx:603-897; spec in SURVEY.md Appendix A).                     */

/* Facet extension (no counterpart in _kernel.pyx; lowers the SurfaceDelegate subclasses of pvtrace/device/lsc.py:22-86
 * and examples/006 Coatings.ipynb cell 3 into data): first facet of `node` whose LOCAL normal equals nl within its
 * tolerance and whose region -- an open box of the local frame, the whole space unless the coating is partial --
 * holds the local point lp. */
static int find_facet(const pvt_scene_t* S, int node, const double* nl, const double* lp) {
  if (S->n_facets <= 0 || !S->facet_count) return -1;
  int f0 = S->facet_start[node], f1 = f0 + S->facet_count[node];
  for (int f = f0; f < f1; ++f) {
    const double* fn = S->facet_normal + 3 * f;
    double tol = S->facet_atol[f];
    if (!(fabs(fn[0] - nl[0]) <= tol && fabs(fn[1] - nl[1]) <= tol && fabs(fn[2] - nl[2]) <= tol)) continue;
    if (S->facet_region) {
      const double* g = S->facet_region + 6 * f;
      if (!(lp[0] > g[0] && lp[1] > g[1] && lp[2] > g[2] && lp[0] < g[3] && lp[1] < g[4] && lp[2] < g[5])) continue;
    }
    return f;
  }
  return -1;
}

/* What the surface of `hit` does to a ray that has just reached it at world point pos (:848-895 + the facet
 * extension): local point, normals, incidence angle, reflectivity, and how to reflect / transmit. */
typedef struct {
  double lp[3], nl[3], nw[3], nf[3], angle, R, n1, n2;
  int fresnel, straight, lambert;
} surface_t;

static void surface_setup(const pvt_scene_t* S, int hit, int container, int adjacent, const double* pos, const double* dir,
                          double wl, surface_t* g) {
  xform_point(S->world_to_local + 16 * hit, pos, g->lp);
  primitive_normal(S->geom_type[hit], S->geom_params + 4 * hit, g->lp, g->nl);
  xform_vector(S->local_to_world + 16 * hit, g->nl, g->nw);
  for (int i = 0; i < 3; ++i) g->nf[i] = g->nw[i];
  if (dot3(g->nf, dir) < 0.0) for (int i = 0; i < 3; ++i) g->nf[i] = -g->nf[i];
  double c = dot3(g->nf, dir);
  if (c > 1.0) c = 1.0; else if (c < -1.0) c = -1.0;
  g->angle = acos(c);
  g->fresnel = S->surface_type[hit] == PVT_SURF_FRESNEL;
  g->n1 = S->refractive_index[container]; g->n2 = S->refractive_index[adjacent];
  g->R = 0.0; g->straight = 0; g->lambert = 0;
  int fixed = 0;
  const int facet = find_facet(S, hit, g->nl, g->lp);
  if (facet >= 0) {
    g->straight = (S->facet_flags[facet] & PVT_FACET_TRANSMIT_STRAIGHT) != 0;
    g->lambert = (S->facet_flags[facet] & PVT_FACET_REFLECT_LAMBERTIAN) != 0;
    double fr = S->facet_reflectivity[facet];
    if (S->facet_refl_n && S->refl_x && S->facet_refl_n[facet] > 0) { /* coating with a reflectivity spectrum */
      int rs = S->facet_refl_start[facet];
      fr = interp_clamped(wl, S->refl_x + rs, S->refl_y + rs, S->facet_refl_n[facet]);
      fr = fr < 0.0 ? 0.0 : (fr > 1.0 ? 1.0 : fr);
    }
    if (fr >= 0.0) { g->R = fr; fixed = 1; }
    /* where no refracted ray exists a facet that transmits by refraction reflects, whatever it states */
    if (fixed && g->fresnel && !g->straight && g->n2 < g->n1 && sqrt(fmax(1.0 - c * c, 0.0)) * (g->n1 / g->n2) > 1.0) g->R = 1.0;
  }
  if (!fixed && g->fresnel) g->R = fresnel_R(g->angle, g->n1, g->n2);
}

/* direction after the reflect / transmit decision; (p1, p2) are the Lambertian uniforms */
static void surface_apply(const surface_t* g, int reflected, const double* dir, double p1, double p2, double* out);

/* Lambertian direction about unit vector n (basis chosen so that n = +z reproduces lambertian(), utils.py:173-186) */
static void lambert_about(const double* n, double p1, double p2, double* out) {
  double loc[3];
  polar_dir(asin(sqrt(p1)), 2.0 * M_PI * p2, loc);
  double t1[3], t2[3];
  if (n[2] < -0.9999999) {
    t1[0] = 0.0; t1[1] = -1.0; t1[2] = 0.0;
    t2[0] = -1.0; t2[1] = 0.0; t2[2] = 0.0;
  } else {
    double a = 1.0 / (1.0 + n[2]);
    double b = -n[0] * n[1] * a;
    t1[0] = 1.0 - n[0] * n[0] * a; t1[1] = b; t1[2] = -n[0];
    t2[0] = b; t2[1] = 1.0 - n[1] * n[1] * a; t2[2] = -n[1];
  }
  for (int i = 0; i < 3; ++i) out[i] = loc[0] * t1[i] + loc[1] * t2[i] + loc[2] * n[i];
}

static void surface_apply(const surface_t* g, int reflected, const double* dir, double p1, double p2, double* out) {
  if (reflected) {
    if (g->lambert) {
      double back[3] = {-g->nf[0], -g->nf[1], -g->nf[2]}; /* hemisphere the ray arrived from */
      lambert_about(back, p1, p2, out);
    } else {
      mirror_dir(dir, g->nw, out);
    }
  } else if (g->fresnel && !g->straight) {
    snell_dir(dir, g->nf, g->n1, g->n2, out);
  } else {
    out[0] = dir[0]; out[1] = dir[1]; out[2] = dir[2];
  }
}

static int trace_photon(const pvt_scene_t* S, const pvt_params_t* P, log_t* L, acc_t* A, double* pos, double* dir,
                        double wl, uint64_t index, int64_t* steps_out) {
  rng_t rng;
  rng_init(&rng, P->rng_mode, P->seed, index);
  unsigned char seen[PVT_MAX_RECORDERS];
  memset(seen, 0, (size_t)(S->n_recorders > 0 ? S->n_recorders : 1));
  double travelled = 0.0, duration = 0.0;
  int source = -1, count = 0;
  int64_t steps = 0;
  const int have_rec = S->n_recorders > 0;

  log_event(L, PVT_EV_GENERATE, -1, -1, -1, -1, source, pos, dir, NULL, wl, travelled, duration);

  for (;;) {
    ++count;
    rng.step = (uint32_t)count;
    /* event budget: leave room for the KILL record (:658-663), sampled rays only */
    if (L->base >= 0 && L->n >= L->max_events - 1) {
      log_event(L, PVT_EV_KILL, -1, -1, -1, -1, source, pos, dir, NULL, wl, travelled, duration);
      break;
    }
    ++steps;
    nearest_t nh;
    nearest_surface(S, pos, dir, &nh);
    if (nh.nhits == 0) break;
    const int hit = nh.hit, container = nh.container, adjacent = nh.adjacent;
    const double t0 = nh.t0;

    if (count > P->maxsteps) { /* :716-723 */
      log_event(L, PVT_EV_KILL, -1, container, -1, -1, source, pos, dir, NULL, wl, travelled, duration);
      if (have_rec) {
        double lp[3];
        xform_point(S->world_to_local + 16 * container, pos, lp);
        tally_event(S, A, PVT_REC_KILLED, container, seen, NULL, lp, 0.0, wl, travelled, duration);
      }
      break;
    }

    const double n_container = S->refractive_index[container];

    if (hit == S->root_id) { /* leave the scene, :728-744 */
      for (int i = 0; i < 3; ++i) pos[i] = pos[i] + dir[i] * t0;
      travelled += t0;
      duration += t0 * n_container / C_CM_PER_S;
      log_event(L, PVT_EV_EXIT, hit, container, adjacent, -1, source, pos, dir, NULL, wl, travelled, duration);
      if (have_rec) {
        double lp[3], nl[3], nw[3];
        xform_point(S->world_to_local + 16 * hit, pos, lp);
        primitive_normal(S->geom_type[hit], S->geom_params + 4 * hit, lp, nl);
        xform_vector(S->local_to_world + 16 * hit, nl, nw);
        double c = fabs(dot3(nw, dir));
        if (c > 1.0) c = 1.0;
        tally_event(S, A, PVT_REC_EXIT, hit, seen, nw, lp, acos(c), wl, travelled, duration);
      }
      break;
    }

    /* Beer-Lambert free path in the container (material.py:17-47 == :746-760) */
    const int c0 = S->comp_start[container], cn = S->comp_count[container];
    double alpha = 0.0;
    for (int k = 0; k < cn; ++k) {
      int c = c0 + k;
      alpha += interp_clamped(wl, S->abs_x + S->comp_abs_start[c], S->abs_y + S->comp_abs_start[c], S->comp_abs_n[c]);
    }
    double depth = INFINITY;
    if (alpha > ALPHA_ZERO) depth = -log(1.0 - rng_draw(&rng, BLOCK_PATH, 0)) / alpha;

    if (depth < t0) { /* absorbed in the volume, :762-832 */
      for (int i = 0; i < 3; ++i) pos[i] = pos[i] + dir[i] * depth;
      travelled += depth;
      duration += depth * n_container / C_CM_PER_S;

      double target = rng_draw(&rng, BLOCK_ABSORB, 0) * alpha, running = 0.0;
      int comp = c0;
      for (int k = 0; k < cn; ++k) {
        int c = c0 + k;
        running += interp_clamped(wl, S->abs_x + S->comp_abs_start[c], S->abs_y + S->comp_abs_start[c], S->comp_abs_n[c]);
        if (target <= running) { comp = c; break; }
      }
      log_event(L, PVT_EV_ABSORB, -1, container, -1, comp, source, pos, dir, NULL, wl, travelled, duration);

      const int ctype = S->comp_type[comp];
      if ((ctype == PVT_COMP_SCATTERER || ctype == PVT_COMP_LUMINOPHORE) &&
          rng_draw(&rng, BLOCK_ABSORB, 1) < S->comp_qy[comp]) {
        double nd[3];
        double g1 = rng_draw(&rng, BLOCK_PHASE, 0), g2 = rng_draw(&rng, BLOCK_PHASE, 1);
        phase_dir(S->comp_phase_type[comp], S->comp_phase_param[comp], g1, g2, nd);
        dir[0] = nd[0]; dir[1] = nd[1]; dir[2] = nd[2];
        source = comp;
        if (ctype == PVT_COMP_LUMINOPHORE) { /* component.py:381-440 == :795-812 */
          const double* ex = S->ems_x + S->comp_ems_start[comp];
          const double* ec = S->ems_cdf + S->comp_ems_start[comp];
          int en = S->comp_ems_n[comp];
          double p1 = 0.0;
          if (P->emit_method != PVT_EMIT_FULL) {
            double nm = wl;
            if (P->emit_method == PVT_EMIT_KT) {
              double ev = 1240.0 / nm + 1.5 * KB_EV * 300.0;
              nm = 1240.0 / ev;
            }
            p1 = interp_clamped(nm, ex, ec, en);
          }
          double gamma = p1 + (1.0 - p1) * rng_draw(&rng, BLOCK_EMIT, 0);
          wl = interp_clamped(gamma, ec, ex, en);
          if (S->comp_tau_rad[comp] > 0.0) duration += -log(1.0 - rng_draw(&rng, BLOCK_EMIT, 1)) * S->comp_tau_rad[comp];
          log_event(L, PVT_EV_EMIT, -1, container, -1, comp, source, pos, dir, NULL, wl, travelled, duration);
        } else {
          log_event(L, PVT_EV_SCATTER, -1, container, -1, comp, source, pos, dir, NULL, wl, travelled, duration);
        }
        continue;
      }
      if (S->comp_tau_nr[comp] > 0.0) duration += -log(1.0 - rng_draw(&rng, BLOCK_EMIT, 1)) * S->comp_tau_nr[comp];
      int sel;
      if (ctype == PVT_COMP_REACTOR) {
        log_event(L, PVT_EV_REACT, -1, container, -1, comp, source, pos, dir, NULL, wl, travelled, duration);
        sel = PVT_REC_REACTED;
      } else {
        log_event(L, PVT_EV_NONRADIATIVE, -1, container, -1, comp, source, pos, dir, NULL, wl, travelled, duration);
        sel = PVT_REC_LOST;
      }
      if (have_rec) {
        double lp[3];
        xform_point(S->world_to_local + 16 * container, pos, lp);
        tally_event(S, A, sel, container, seen, NULL, lp, 0.0, wl, travelled, duration);
      }
      break;
    }

    /* reach the surface, :834-895 */
    for (int i = 0; i < 3; ++i) pos[i] = pos[i] + dir[i] * t0;
    travelled += t0;
    duration += t0 * n_container / C_CM_PER_S;
    if (adjacent < 0) {
      log_event(L, PVT_EV_KILL, hit, container, -1, -1, source, pos, dir, NULL, wl, travelled, duration);
      break;
    }
    surface_t g;
    surface_setup(S, hit, container, adjacent, pos, dir, wl, &g);
    double u = 1.0;
    if (g.R > 0.0) u = rng_draw(&rng, BLOCK_PATH, 1); /* surface.py:231-240: no draw when R == 0 */
    double nd[3];
    if (u < g.R) {
      double p1 = 0.0, p2 = 0.0;
      if (g.lambert) { p1 = rng_draw(&rng, BLOCK_LAMBERT, 0); p2 = rng_draw(&rng, BLOCK_LAMBERT, 1); }
      surface_apply(&g, 1, dir, p1, p2, nd);
      dir[0] = nd[0]; dir[1] = nd[1]; dir[2] = nd[2];
      log_event(L, PVT_EV_REFLECT, hit, container, adjacent, -1, source, pos, dir, g.nw, wl, travelled, duration);
      if (have_rec && container != hit)
        tally_event(S, A, PVT_REC_REFLECTED, hit, seen, g.nw, g.lp, g.angle, wl, travelled, duration);
    } else {
      surface_apply(&g, 0, dir, 0.0, 0.0, nd);
      dir[0] = nd[0]; dir[1] = nd[1]; dir[2] = nd[2];
      log_event(L, PVT_EV_TRANSMIT, hit, container, adjacent, -1, source, pos, dir, g.nw, wl, travelled, duration);
      if (have_rec)
        tally_event(S, A, container == hit ? PVT_REC_ESCAPING : PVT_REC_ENTERING, hit, seen, g.nw, g.lp, g.angle, wl,
                    travelled, duration);
    }
  }
  *steps_out = steps;
  return L->n;
}

/* ------------------------------------------------------------------------------------------------
 * Seeded emission of built-in light delegates (emit.py:22-134; scene.py:141-151 for the round robin).
 * Draw k of Philox stream 1 of photon `index` under key `seed`: k=0 wavelength, k=1..3 position, k=4,5 direction.               */

static void emit_one(const pvt_emit_t* E, uint64_t seed, int64_t index, double* pos, double* dir, double* wl) {
  const uint64_t id = (uint64_t)index;
  int l = (int)(index % E->n_lights);
  double lp[3] = {0.0, 0.0, 0.0}, ld[3] = {0.0, 0.0, 1.0};
  if (E->wl_kind[l] == PVT_LWL_SPECTRUM) {
    double u = philox_uniform_at(seed, id, 1u, 0u);
    *wl = interp_clamped(u, E->wl_cdf + E->wl_start[l], E->wl_x + E->wl_start[l], E->wl_n[l]);
  } else {
    *wl = E->wl_param[l];
  }
  const double* pp = E->pos_param + 3 * l;
  switch (E->pos_kind[l]) {
    case PVT_LPOS_RECT:
      lp[0] = -pp[0] + (pp[0] - -pp[0]) * philox_uniform_at(seed, id, 1u, 1u);
      lp[1] = -pp[1] + (pp[1] - -pp[1]) * philox_uniform_at(seed, id, 1u, 2u);
      break;
    case PVT_LPOS_CIRCLE: {
      double ang = 2.0 * M_PI * philox_uniform_at(seed, id, 1u, 1u);
      double r = sqrt(philox_uniform_at(seed, id, 1u, 2u)) * pp[0];
      lp[0] = r * cos(ang); lp[1] = r * sin(ang);
    } break;
    case PVT_LPOS_CUBE:
      for (int i = 0; i < 3; ++i) lp[i] = -pp[i] + (pp[i] - -pp[i]) * philox_uniform_at(seed, id, 1u, 1u + i);
      break;
    default: break;
  }
  double u0 = philox_uniform_at(seed, id, 1u, 4u), u1 = philox_uniform_at(seed, id, 1u, 5u);
  double prm = E->dir_param[l];
  int kind = E->dir_kind[l];
  if (kind == PVT_LDIR_HG && fabs(prm) < 1e-12) kind = PVT_LDIR_ISOTROPIC;
  switch (kind) {
    case PVT_LDIR_CONE: polar_dir(asin(sqrt(u0) * sin(prm)), 2.0 * M_PI * u1, ld); break;
    case PVT_LDIR_ISOTROPIC: polar_dir(acos(2.0 * u1 - 1.0), 2.0 * M_PI * u0, ld); break;
    case PVT_LDIR_LAMBERTIAN: polar_dir(asin(sqrt(u0)), 2.0 * M_PI * u1, ld); break;
    case PVT_LDIR_HG: {
      double s = 2.0 * u0 - 1.0;
      double f = (1.0 - prm * prm) / (1.0 + prm * s);
      double mu = (1.0 + prm * prm - f * f) / (2.0 * prm);
      polar_dir(acos(mu), 2.0 * M_PI * u1, ld);
    } break;
    default: break;
  }
  xform_point(E->light_to_world + 16 * l, lp, pos);
  xform_vector(E->light_to_world + 16 * l, ld, dir);
}

/* ================================================================================================
 * Exported entry points (ctypes).  Mirrors of the product ABI, prefixed pvt_oracle_.                 */

int pvt_oracle_emit_bundle(const pvt_emit_t* E, double* pos, double* dir, double* wl, int64_t n,
                           int64_t first_index, uint64_t seed) {
  if (!E || E->n_lights <= 0) return 1;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    emit_one(E, seed, first_index + i, pos + 3 * i, dir + 3 * i, wl + i);
  return 0;
}

int pvt_oracle_trace_bundle(const pvt_scene_t* S, const pvt_emit_t* E, const double* positions,
                            const double* directions, const double* wavelengths, const pvt_params_t* P,
                            pvt_out_t* out, int num_threads) {
  if (S->n_nodes > PVT_MAX_NODES || S->n_recorders > PVT_MAX_RECORDERS) return 1;
  const int64_t n = P->n;
  const int R = S->n_recorders, B = S->total_bins;
  int nthr = num_threads > 0 ? num_threads : 1;
#ifndef _OPENMP
  nthr = 1;
#endif
  /* per-thread tally slabs merged at the end, like _kernel.pyx:1019-1032,1097-1102 */
  int64_t* t_distinct = (int64_t*)calloc((size_t)nthr * (R > 0 ? R : 1), sizeof(int64_t));
  int64_t* t_cross = (int64_t*)calloc((size_t)nthr * (R > 0 ? R : 1), sizeof(int64_t));
  double* t_sums = (double*)calloc((size_t)nthr * (R > 0 ? R : 1) * 8, sizeof(double));
  int64_t* t_bins = (int64_t*)calloc((size_t)nthr * (B > 0 ? B : 1), sizeof(int64_t));
  int64_t* t_steps = (int64_t*)calloc((size_t)nthr, sizeof(int64_t));
  int64_t* t_events = (int64_t*)calloc((size_t)nthr, sizeof(int64_t));
  if (!t_distinct || !t_cross || !t_sums || !t_bins || !t_steps || !t_events) return 2;

#pragma omp parallel for schedule(dynamic, 64) num_threads(nthr)
  for (int64_t i = 0; i < n; ++i) {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    acc_t A = {t_distinct + (size_t)tid * R, t_cross + (size_t)tid * R, t_sums + (size_t)tid * R * 8,
               t_bins + (size_t)tid * B};
    log_t L = {out, -1, P->max_events, 0};
    if (P->record_every > 0 && i % P->record_every == 0) L.base = (i / P->record_every) * (int64_t)P->max_events;
    uint64_t id = (uint64_t)P->first_index + (uint64_t)i; /* photon index within the run */
    double pos[3], dir[3], wl;
    if (positions) {
      for (int k = 0; k < 3; ++k) { pos[k] = positions[3 * i + k]; dir[k] = directions[3 * i + k]; }
      wl = wavelengths[i];
    } else {
      emit_one(E, P->seed, P->first_index + i, pos, dir, &wl);
    }
    int64_t steps = 0;
    int nev = trace_photon(S, P, &L, &A, pos, dir, wl, id, &steps);
    t_steps[tid] += steps;
    t_events[tid] += nev;
    if (L.base >= 0) out->counts[i / P->record_every] = nev;
  }

  for (int r = 0; r < R; ++r) { out->rec_distinct[r] = 0; out->rec_crossings[r] = 0; }
  for (int k = 0; k < R * 8; ++k) out->rec_sums[k] = 0.0;
  for (int b = 0; b < B; ++b) out->rec_bins[b] = 0;
  int64_t steps = 0;
  for (int t = 0; t < nthr; ++t) {
    for (int r = 0; r < R; ++r) {
      out->rec_distinct[r] += t_distinct[(size_t)t * R + r];
      out->rec_crossings[r] += t_cross[(size_t)t * R + r];
    }
    for (int k = 0; k < R * 8; ++k) out->rec_sums[k] += t_sums[(size_t)t * R * 8 + k];
    for (int b = 0; b < B; ++b) out->rec_bins[b] += t_bins[(size_t)t * B + b];
    steps += t_steps[t];
  }
  if (out->stats) {
    memset(out->stats, 0, sizeof(int64_t) * PVT_NSTATS);
    out->stats[PVT_STAT_STEPS] = steps;
    out->stats[PVT_STAT_RAYS] = n;
  }
  free(t_distinct); free(t_cross); free(t_sums); free(t_bins); free(t_steps); free(t_events);
  return 0;
}

int pvt_oracle_intersect_bundle(const pvt_scene_t* S, const double* positions, const double* directions, int64_t n,
                                double* t0, int32_t* hit, int32_t* container, int32_t* adjacent) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    nearest_t nh;
    nearest_surface(S, positions + 3 * i, directions + 3 * i, &nh);
    t0[i] = nh.nhits ? nh.t0 : INFINITY;
    hit[i] = nh.nhits ? nh.hit : -1;
    container[i] = nh.container;
    adjacent[i] = nh.adjacent;
  }
  return 0;
}

/* One surface interaction on its own (pins the facet extension against the reference's SurfaceDelegate classes called
 * directly, tests/test_reference_pins.py): `pos` lies on the surface of `hit`; out_R = reflectivity, out_reflect /
 * out_transmit = the two candidate directions (Lambertian uniforms p1, p2). */
int pvt_oracle_surface_event(const pvt_scene_t* S, int64_t n, const int32_t* hit, const int32_t* container,
                             const int32_t* adjacent, const double* pos, const double* dir, const double* wl,
                             const double* p1, const double* p2, double* out_R, double* out_reflect, double* out_transmit) {
  for (int64_t i = 0; i < n; ++i) {
    surface_t g;
    surface_setup(S, hit[i], container[i], adjacent[i], pos + 3 * i, dir + 3 * i, wl[i], &g);
    out_R[i] = g.R;
    surface_apply(&g, 1, dir + 3 * i, p1 ? p1[i] : 0.5, p2 ? p2[i] : 0.5, out_reflect + 3 * i);
    if (g.R < 1.0) surface_apply(&g, 0, dir + 3 * i, 0.0, 0.0, out_transmit + 3 * i);
    else for (int k = 0; k < 3; ++k) out_transmit[3 * i + k] = NAN;
  }
  return 0;
}

/* known-answer helpers, one-to-one with pvt_test_* of the product ABI */
int pvt_oracle_fresnel_reflectivity(int64_t n, const double* angle, const double* n1, const double* n2, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = fresnel_R(angle[i], n1[i], n2[i]);
  return 0;
}
int pvt_oracle_specular_reflect(int64_t n, const double* d, const double* nrm, double* out) {
  for (int64_t i = 0; i < n; ++i) mirror_dir(d + 3 * i, nrm + 3 * i, out + 3 * i);
  return 0;
}
int pvt_oracle_fresnel_refract(int64_t n, const double* d, const double* nrm, const double* n1, const double* n2,
                               double* out) {
  for (int64_t i = 0; i < n; ++i) {
    double nf[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
    if (dot3(nf, d + 3 * i) < 0.0) { nf[0] = -nf[0]; nf[1] = -nf[1]; nf[2] = -nf[2]; }
    snell_dir(d + 3 * i, nf, n1[i], n2[i], out + 3 * i);
  }
  return 0;
}
int pvt_oracle_intersect(int64_t n, const int32_t* gtype, const double* params, const double* o, const double* d,
                         int32_t* nhit, double* ts) {
  for (int64_t i = 0; i < n; ++i) {
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    nhit[i] = hit_primitive(gtype[i], params + 4 * i, o + 3 * i, d + 3 * i, t);
    for (int k = 0; k < 4; ++k) ts[4 * i + k] = k < nhit[i] ? t[k] : 0.0;
  }
  return 0;
}
int pvt_oracle_local_normal(int64_t n, const int32_t* gtype, const double* params, const double* p, double* out) {
  for (int64_t i = 0; i < n; ++i) primitive_normal(gtype[i], params + 4 * i, p + 3 * i, out + 3 * i);
  return 0;
}
int pvt_oracle_interp(int64_t n, const double* x, int32_t m, const double* xs, const double* ys, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = interp_clamped(x[i], xs, ys, m);
  return 0;
}
int pvt_oracle_rng_uniform(int64_t n_rays, int32_t n_draws, uint64_t seed, int64_t first_index, int32_t rng_mode,
                           double* out) {
  for (int64_t i = 0; i < n_rays; ++i) {
    rng_t r;
    rng_init(&r, rng_mode, seed, (uint64_t)first_index + (uint64_t)i);
    for (int k = 0; k < n_draws; ++k) out[i * n_draws + k] = rng_next(&r);
  }
  return 0;
}
int pvt_oracle_sample_phase(int64_t n, int32_t phase_type, double phase_param, uint64_t seed, int32_t rng_mode,
                            double* out) {
  for (int64_t i = 0; i < n; ++i) {
    rng_t r;
    rng_init(&r, rng_mode, seed, (uint64_t)i);
    double g1 = rng_next(&r), g2 = rng_next(&r);
    phase_dir(phase_type, phase_param, g1, g2, out + 3 * i);
  }
  return 0;
}
/* raw Philox block for the published known-answer vectors (Random123 kat_vectors) */
void pvt_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox4x32_10(c, key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}
