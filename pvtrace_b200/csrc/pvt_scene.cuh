// pvt_scene.cuh -- the device-resident scene: ONE contiguous blob of 8-byte words that every CTA stages into
// shared memory with a single bulk async copy (cp.async.bulk, the TMA engine's 1-D path) before tracing.
//
// Host side: pack_scene() turns the flat tables of include/pvtrace_b200.h (pvt_scene_t / pvt_emit_t, themselves
// the reference's CompiledScene contract, pvtrace/engine/compiler.py:57-204) into the blob.  Per-entity records
// (node / component / recorder / histogram / facet / light) are arrays-of-structs because every lane of a warp
// reads the SAME record at the same time (shared-memory broadcast); spectra stay as plain arrays.
//
// Word layout (all offsets in 8-byte words from the start of the blob):
//   [0, kHeaderWords)  header: int32 pairs, see Header
//   nodes       n_nodes      x kNodeWords
//   components  n_components x kCompWords
//   abs_x, abs_y, ems_x, ems_cdf  (spectra knots)
//   recorders   n_recorders  x kRecWords
//   hists       n_hists      x kHistWords
//   facets      n_facets     x kFacetWords
//   lights      n_lights     x kLightWords, wl_x, wl_cdf
//   rec_index   n_nodes x kRecSelectors (start, count) int pairs into rec_list: the recorders attached to
//               (node, selector), so a tally looks at its handful of candidates instead of every recorder
//   rec_list    n_recorders int32 recorder indices grouped by (node, selector), original order within a group
//   face_index  n_nodes x kRecSelectors x 6 (start, count) int pairs into face_list: for BOX nodes, the recorders
//               of (node, selector) that match a surface event on local face f (-x +x -y +y -z +z) -- the facet
//               test against the face's world normal is done once here instead of per event; count < 0 marks a
//               node whose faces cannot be pre-resolved (not a box, or a facet within 1e-9 of its tolerance)
//   face_list   int32 recorder indices
//   refl_x, refl_y  knots of the wavelength-tabulated facet reflectivities (coatings)
//   ems_guide   per emitting component, kGuideBuckets + 1 uint16: guide[b] = last knot i with ems_cdf[i] <= b / kGuideBuckets
//               (0 if none), so the inverse-CDF lookup bisects a bracket of a knot or two instead of the whole table
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../include/pvtrace_b200.h"

namespace pvt {

struct Header {
  int32_t n_nodes, root_id, n_components, n_recorders, n_hists, total_bins, n_facets, n_lights;
  int32_t off_nodes, off_comps, off_abs_x, off_abs_y, off_ems_x, off_ems_cdf, off_recs, off_hists;
  int32_t off_facets, off_lights, off_wl_x, off_wl_cdf, total_words, off_rec_index, off_rec_list, off_face_index;
  int32_t off_face_list, off_ems_guide, off_refl_x, off_refl_y;
};
constexpr int kHeaderWords = sizeof(Header) / 8;
static_assert(sizeof(Header) % 16 == 0, "header must keep 16-byte alignment for the bulk copy");

// node record: w2l rows 0-2 (12) | l2w rows 0-2 (12) | params (4) | n (1) | ints: geom,surf | comp_start,comp_count |
// facet_start,facet_count | n / c (seconds per cm) | ints: aligned (rotation part of w2l is exactly the identity), pad |
// half (3): 0.5 * params, exact -- the box slabs sit at -half and +half | pad
constexpr int kNodeW2L = 0, kNodeL2W = 12, kNodeParams = 24, kNodeIndex = 28, kNodeInts = 29, kNodeSlowness = 32,
              kNodeHalf = 34, kNodeWords = 38;
// component record: qy, tau_rad, tau_nr, phase_param | ints: type,phase_type | abs_start,abs_n | ems_start,ems_n |
// abs_inv_dx, ems_inv_dx (1/spacing of the x grid when it is uniform enough for interp_hinted, else 0) |
// ints: guide_start (uint16 index into ems_guide), has_guide
constexpr int kGuideBuckets = 256;
constexpr int kCompQy = 0, kCompTauRad = 1, kCompTauNr = 2, kCompPhaseParam = 3, kCompInts = 4, kCompAbsInvDx = 7,
              kCompEmsInvDx = 8, kCompWords = 10;
// recorder record: facet xyz, atol | ints: node,event | has_facet,hist_start | hist_n,pad | pad
constexpr int kRecFacet = 0, kRecAtol = 3, kRecInts = 4, kRecWords = 8;
// histogram record: lo_a, hi_a, lo_b, hi_b | ints: prop_a,prop_b | na,nb | offset,pad | pad
constexpr int kHistLoA = 0, kHistHiA = 1, kHistLoB = 2, kHistHiB = 3, kHistInts = 4, kHistWords = 8;
// facet record: normal xyz, atol, reflectivity | ints: flags,pad | ints: refl_start,refl_n | pad | region lo xyz, hi xyz | pad pad
constexpr int kFacetNormal = 0, kFacetAtol = 3, kFacetRefl = 4, kFacetInts = 5, kFacetRegion = 8, kFacetWords = 16;
// light record: l2w rows 0-2 (12) | pos_param (3) | dir_param | wl_param | ints: pos_kind,dir_kind | wl_kind,wl_start | wl_n,pad
constexpr int kRecSelectors = 8;  // PVT_REC_* selectors 0..6, padded to 8
constexpr int kLightL2W = 0, kLightPos = 12, kLightDir = 15, kLightWl = 16, kLightInts = 17, kLightSinDir = 20 /* sin(dir_param) */,
              kLightWords = 22;

// 1/dx when xs[0..n) is a uniform ascending grid (every knot within 1e-9 dx of xs[0] + i dx, so multiplying by
// 1/dx stands in for dividing by the local spacing, and a guessed bracket is at most one knot off).  0 otherwise.
inline double uniform_inv_dx(const double* xs, int n) {
  if (n < 3) return 0.0;
  const double dx = (xs[n - 1] - xs[0]) / (n - 1);
  if (!(dx > 0.0)) return 0.0;
  for (int i = 0; i < n; ++i) {
    const double dev = xs[i] - (xs[0] + i * dx);
    if (dev > 1e-9 * dx || dev < -1e-9 * dx) return 0.0;
  }
  return 1.0 / dx;
}

inline void put_ints(std::vector<double>& blob, size_t word, int32_t a, int32_t b) {
  int32_t pair[2] = {a, b};
  memcpy(&blob[word], pair, 8);
}

// Returns the blob; its size is a multiple of 2 words (16 bytes).
inline std::vector<double> pack_scene(const pvt_scene_t& S, const pvt_emit_t* E) {
  Header h;
  memset(&h, 0, sizeof(h));
  h.n_nodes = S.n_nodes; h.root_id = S.root_id; h.n_components = S.n_components; h.n_recorders = S.n_recorders;
  h.n_hists = S.n_hists; h.total_bins = S.total_bins; h.n_facets = S.facet_count ? S.n_facets : 0;
  h.n_lights = E ? E->n_lights : 0;
  int32_t w = kHeaderWords;
  h.off_nodes = w;   w += S.n_nodes * kNodeWords;
  h.off_comps = w;   w += S.n_components * kCompWords;
  h.off_abs_x = w;   w += S.n_abs_knots;
  h.off_abs_y = w;   w += S.n_abs_knots;
  h.off_ems_x = w;   w += S.n_ems_knots;
  h.off_ems_cdf = w; w += S.n_ems_knots;
  h.off_recs = w;    w += S.n_recorders * kRecWords;
  h.off_hists = w;   w += S.n_hists * kHistWords;
  h.off_facets = w;  w += h.n_facets * kFacetWords;
  h.off_lights = w;  w += h.n_lights * kLightWords;
  h.off_wl_x = w;    w += E ? E->n_wl_knots : 0;
  h.off_wl_cdf = w;  w += E ? E->n_wl_knots : 0;
  h.off_rec_index = w; w += S.n_nodes * kRecSelectors;
  h.off_rec_list = w;  w += (S.n_recorders + 1) / 2;
  h.off_face_index = w; w += S.n_nodes * kRecSelectors * 6;
  // worst case every recorder of a node matches all six faces
  h.off_face_list = w; w += (6 * S.n_recorders + 1) / 2;
  h.off_ems_guide = w; w += (S.n_components * (kGuideBuckets + 1) + 3) / 4;
  const int n_refl = (h.n_facets && S.facet_refl_n && S.refl_x && S.refl_y) ? S.n_refl_knots : 0;
  h.off_refl_x = w;  w += n_refl;
  h.off_refl_y = w;  w += n_refl;
  w = (w + 1) & ~1;
  h.total_words = w;

  std::vector<double> blob((size_t)w, 0.0);
  memcpy(blob.data(), &h, sizeof(h));
  for (int i = 0; i < S.n_nodes; ++i) {
    double* r = &blob[h.off_nodes + (size_t)i * kNodeWords];
    memcpy(r + kNodeW2L, S.world_to_local + 16 * i, 12 * sizeof(double));
    memcpy(r + kNodeL2W, S.local_to_world + 16 * i, 12 * sizeof(double));
    memcpy(r + kNodeParams, S.geom_params + 4 * i, 4 * sizeof(double));
    r[kNodeIndex] = S.refractive_index[i];
    const size_t iw = h.off_nodes + (size_t)i * kNodeWords + kNodeInts;
    put_ints(blob, iw, S.geom_type[i], S.surface_type[i]);
    put_ints(blob, iw + 1, S.comp_start[i], S.comp_count[i]);
    put_ints(blob, iw + 2, h.n_facets ? S.facet_start[i] : 0, h.n_facets ? S.facet_count[i] : 0);
    r[kNodeSlowness] = S.refractive_index[i] / 2.99792458e10;
    const double* m = S.world_to_local + 16 * i;
    const bool aligned = m[0] == 1.0 && m[1] == 0.0 && m[2] == 0.0 && m[4] == 0.0 && m[5] == 1.0 && m[6] == 0.0 &&
                         m[8] == 0.0 && m[9] == 0.0 && m[10] == 1.0;
    put_ints(blob, iw + 4, aligned ? 1 : 0, 0);
    for (int k = 0; k < 3; ++k) r[kNodeHalf + k] = 0.5 * S.geom_params[4 * i + k];
  }
  for (int c = 0; c < S.n_components; ++c) {
    double* r = &blob[h.off_comps + (size_t)c * kCompWords];
    r[kCompQy] = S.comp_qy[c]; r[kCompTauRad] = S.comp_tau_rad[c]; r[kCompTauNr] = S.comp_tau_nr[c];
    r[kCompPhaseParam] = S.comp_phase_param[c];
    const size_t iw = h.off_comps + (size_t)c * kCompWords + kCompInts;
    put_ints(blob, iw, S.comp_type[c], S.comp_phase_type[c]);
    put_ints(blob, iw + 1, S.comp_abs_start[c], S.comp_abs_n[c]);
    put_ints(blob, iw + 2, S.comp_ems_start[c], S.comp_ems_n[c]);
    r[kCompAbsInvDx] = uniform_inv_dx(S.abs_x + S.comp_abs_start[c], S.comp_abs_n[c]);
    r[kCompEmsInvDx] = S.comp_ems_n[c] > 0 ? uniform_inv_dx(S.ems_x + S.comp_ems_start[c], S.comp_ems_n[c]) : 0.0;
    {  // guide table of the inverse CDF: valid for a non-decreasing cdf that ends at or below 1
      const int n = S.comp_ems_n[c];
      const double* cdf = S.ems_cdf + S.comp_ems_start[c];
      bool ok = n >= 2 && n < 65535 && cdf[n - 1] <= 1.0;
      for (int i = 1; i < n && ok; ++i) ok = cdf[i] >= cdf[i - 1];
      uint16_t* guide = reinterpret_cast<uint16_t*>(&blob[h.off_ems_guide]) + (size_t)c * (kGuideBuckets + 1);
      if (ok) {
        int i = 0;
        for (int b = 0; b <= kGuideBuckets; ++b) {
          const double edge = (double)b / kGuideBuckets;
          while (i + 1 < n && cdf[i + 1] <= edge) ++i;
          guide[b] = (uint16_t)((cdf[i] <= edge) ? i : 0);
        }
      }
      put_ints(blob, iw + 5, c * (kGuideBuckets + 1), ok ? 1 : 0);
    }
  }
  if (S.n_abs_knots) {
    memcpy(&blob[h.off_abs_x], S.abs_x, S.n_abs_knots * sizeof(double));
    memcpy(&blob[h.off_abs_y], S.abs_y, S.n_abs_knots * sizeof(double));
  }
  if (S.n_ems_knots) {
    memcpy(&blob[h.off_ems_x], S.ems_x, S.n_ems_knots * sizeof(double));
    memcpy(&blob[h.off_ems_cdf], S.ems_cdf, S.n_ems_knots * sizeof(double));
  }
  for (int r = 0; r < S.n_recorders; ++r) {
    double* q = &blob[h.off_recs + (size_t)r * kRecWords];
    q[0] = S.rec_facet[3 * r]; q[1] = S.rec_facet[3 * r + 1]; q[2] = S.rec_facet[3 * r + 2];
    q[kRecAtol] = S.rec_atol[r];
    const size_t iw = h.off_recs + (size_t)r * kRecWords + kRecInts;
    put_ints(blob, iw, S.rec_node[r], S.rec_event[r]);
    put_ints(blob, iw + 1, S.rec_has_facet[r], S.rec_hist_start[r]);
    put_ints(blob, iw + 2, S.rec_hist_n[r], 0);
  }
  for (int k = 0; k < S.n_hists; ++k) {
    double* q = &blob[h.off_hists + (size_t)k * kHistWords];
    q[kHistLoA] = S.hist_lo_a[k]; q[kHistHiA] = S.hist_hi_a[k]; q[kHistLoB] = S.hist_lo_b[k]; q[kHistHiB] = S.hist_hi_b[k];
    const size_t iw = h.off_hists + (size_t)k * kHistWords + kHistInts;
    put_ints(blob, iw, S.hist_prop_a[k], S.hist_prop_b[k]);
    put_ints(blob, iw + 1, S.hist_na[k], S.hist_nb[k]);
    put_ints(blob, iw + 2, S.hist_offset[k], 0);
  }
  for (int f = 0; f < h.n_facets; ++f) {
    double* q = &blob[h.off_facets + (size_t)f * kFacetWords];
    q[0] = S.facet_normal[3 * f]; q[1] = S.facet_normal[3 * f + 1]; q[2] = S.facet_normal[3 * f + 2];
    q[kFacetAtol] = S.facet_atol[f]; q[kFacetRefl] = S.facet_reflectivity[f];
    put_ints(blob, h.off_facets + (size_t)f * kFacetWords + kFacetInts, S.facet_flags[f], 0);
    put_ints(blob, h.off_facets + (size_t)f * kFacetWords + kFacetInts + 1, n_refl ? S.facet_refl_start[f] : 0,
             n_refl ? S.facet_refl_n[f] : 0);
    for (int k = 0; k < 6; ++k)
      q[kFacetRegion + k] = S.facet_region ? S.facet_region[6 * f + k] : (k < 3 ? -HUGE_VAL : HUGE_VAL);
  }
  if (n_refl) {
    memcpy(&blob[h.off_refl_x], S.refl_x, n_refl * sizeof(double));
    memcpy(&blob[h.off_refl_y], S.refl_y, n_refl * sizeof(double));
  }
  for (int l = 0; l < h.n_lights; ++l) {
    double* q = &blob[h.off_lights + (size_t)l * kLightWords];
    memcpy(q + kLightL2W, E->light_to_world + 16 * l, 12 * sizeof(double));
    q[kLightPos] = E->pos_param[3 * l]; q[kLightPos + 1] = E->pos_param[3 * l + 1]; q[kLightPos + 2] = E->pos_param[3 * l + 2];
    q[kLightDir] = E->dir_param[l]; q[kLightWl] = E->wl_param[l];
    q[kLightSinDir] = sin(E->dir_param[l]);
    const size_t iw = h.off_lights + (size_t)l * kLightWords + kLightInts;
    put_ints(blob, iw, E->pos_kind[l], E->dir_kind[l]);
    put_ints(blob, iw + 1, E->wl_kind[l], E->wl_start[l]);
    put_ints(blob, iw + 2, E->wl_n[l], 0);
  }
  {
    int32_t* list = reinterpret_cast<int32_t*>(&blob[h.off_rec_list]);
    int32_t filled = 0;
    for (int node = 0; node < S.n_nodes; ++node)
      for (int sel = 0; sel < kRecSelectors; ++sel) {
        const int32_t start = filled;
        for (int r = 0; r < S.n_recorders; ++r)
          if (S.rec_node[r] == node && S.rec_event[r] == sel) list[filled++] = r;
        put_ints(blob, h.off_rec_index + (size_t)node * kRecSelectors + sel, start, filled - start);
      }
  }
  {
    int32_t* list = reinterpret_cast<int32_t*>(&blob[h.off_face_list]);
    int32_t filled = 0;
    for (int node = 0; node < S.n_nodes; ++node) {
      const double* m = S.local_to_world + 16 * node;
      bool resolvable = S.geom_type[node] == PVT_GEOM_BOX;
      // a facet whose distance to a face normal is within 1e-9 of its tolerance could go either way on the device
      for (int f = 0; f < 6 && resolvable; ++f) {
        const double sg = (f & 1) ? 1.0 : -1.0;
        const int ax = f >> 1;
        const double nw[3] = {m[ax] * sg, m[4 + ax] * sg, m[8 + ax] * sg};
        for (int r = 0; r < S.n_recorders; ++r) {
          if (S.rec_node[r] != node || !S.rec_has_facet[r]) continue;
          for (int k = 0; k < 3; ++k) {
            const double gap = fabs(S.rec_facet[3 * r + k] - nw[k]) - S.rec_atol[r];
            if (fabs(gap) < 1e-9) resolvable = false;
          }
        }
      }
      for (int sel = 0; sel < kRecSelectors; ++sel)
        for (int f = 0; f < 6; ++f) {
          const size_t word = h.off_face_index + ((size_t)node * kRecSelectors + sel) * 6 + f;
          if (!resolvable) { put_ints(blob, word, 0, -1); continue; }
          const double sg = (f & 1) ? 1.0 : -1.0;
          const int ax = f >> 1;
          const double nw[3] = {m[ax] * sg, m[4 + ax] * sg, m[8 + ax] * sg};
          const int32_t start = filled;
          for (int r = 0; r < S.n_recorders; ++r) {
            if (S.rec_node[r] != node || S.rec_event[r] != sel) continue;
            bool match = true;
            if (S.rec_has_facet[r])
              for (int k = 0; k < 3; ++k)
                if (fabs(S.rec_facet[3 * r + k] - nw[k]) > S.rec_atol[r]) match = false;
            if (match) list[filled++] = r;
          }
          put_ints(blob, word, start, filled - start);
        }
    }
  }
  if (E && E->n_wl_knots) {
    memcpy(&blob[h.off_wl_x], E->wl_x, E->n_wl_knots * sizeof(double));
    memcpy(&blob[h.off_wl_cdf], E->wl_cdf, E->n_wl_knots * sizeof(double));
  }
  return blob;
}

#ifdef __CUDACC__
// Read-only typed view over a blob held in shared (or, for oversized scenes, global) memory.
struct SceneView {
  const double* w;   // blob words
  const Header* h;   // the header: a copy in the kernel's parameter space when the kernel has one (offsets then come
                     // from the constant bank, not from a dependent shared-memory load), else the blob's own
  __device__ __forceinline__ SceneView(const double* words) : w(words), h(reinterpret_cast<const Header*>(words)) {}
  __device__ __forceinline__ SceneView(const double* words, const Header* header) : w(words), h(header) {}
  __device__ __forceinline__ const Header& hdr() const { return *h; }
  __device__ __forceinline__ int ival(int word, int half) const { return reinterpret_cast<const int32_t*>(w)[2 * word + half]; }
  __device__ __forceinline__ const double* node(int i) const { return w + hdr().off_nodes + i * kNodeWords; }
  __device__ __forceinline__ int node_int(int i, int k) const { return ival(hdr().off_nodes + i * kNodeWords + kNodeInts + (k >> 1), k & 1); }
  __device__ __forceinline__ const double* comp(int c) const { return w + hdr().off_comps + c * kCompWords; }
  __device__ __forceinline__ int comp_int(int c, int k) const { return ival(hdr().off_comps + c * kCompWords + kCompInts + (k >> 1), k & 1); }
  __device__ __forceinline__ const uint16_t* ems_guide(int c) const {
    return reinterpret_cast<const uint16_t*>(w + hdr().off_ems_guide) + comp_int(c, 10 /* CI_GUIDE_START */);
  }
  __device__ __forceinline__ const double* rec(int r) const { return w + hdr().off_recs + r * kRecWords; }
  __device__ __forceinline__ int rec_int(int r, int k) const { return ival(hdr().off_recs + r * kRecWords + kRecInts + (k >> 1), k & 1); }
  __device__ __forceinline__ const double* hist(int h) const { return w + hdr().off_hists + h * kHistWords; }
  __device__ __forceinline__ int hist_int(int h, int k) const { return ival(hdr().off_hists + h * kHistWords + kHistInts + (k >> 1), k & 1); }
  __device__ __forceinline__ const double* facet(int f) const { return w + hdr().off_facets + f * kFacetWords; }
  __device__ __forceinline__ int facet_flags(int f) const { return ival(hdr().off_facets + f * kFacetWords + kFacetInts, 0); }
  __device__ __forceinline__ int facet_refl_start(int f) const { return ival(hdr().off_facets + f * kFacetWords + kFacetInts + 1, 0); }
  __device__ __forceinline__ int facet_refl_n(int f) const { return ival(hdr().off_facets + f * kFacetWords + kFacetInts + 1, 1); }
  // recorders attached to (node, selector): rec_candidate(start + k), k < count
  __device__ __forceinline__ void rec_range(int node, int sel, int& start, int& count) const {
    const int word = hdr().off_rec_index + node * kRecSelectors + sel;
    start = ival(word, 0); count = ival(word, 1);
  }
  // does any recorder listen to (node, selector)?  Events nobody listens to are not even handed to the tally
  __device__ __forceinline__ bool has_recorders(int node, int sel) const {
    return ival(hdr().off_rec_index + node * kRecSelectors + sel, 1) > 0;
  }
  __device__ __forceinline__ int rec_candidate(int k) const { return reinterpret_cast<const int32_t*>(w + hdr().off_rec_list)[k]; }
  // recorders matching a surface event on local face `face` of a box node; count < 0: not pre-resolved
  __device__ __forceinline__ void face_range(int node, int sel, int face, int& start, int& count) const {
    const int word = hdr().off_face_index + (node * kRecSelectors + sel) * 6 + face;
    start = ival(word, 0); count = ival(word, 1);
  }
  __device__ __forceinline__ int face_candidate(int k) const { return reinterpret_cast<const int32_t*>(w + hdr().off_face_list)[k]; }
  __device__ __forceinline__ const double* light(int l) const { return w + hdr().off_lights + l * kLightWords; }
  __device__ __forceinline__ int light_int(int l, int k) const { return ival(hdr().off_lights + l * kLightWords + kLightInts + (k >> 1), k & 1); }
};
// int slots of the records
enum { NI_GEOM = 0, NI_SURF = 1, NI_COMP_START = 2, NI_COMP_COUNT = 3, NI_FACET_START = 4, NI_FACET_COUNT = 5,
       NI_ALIGNED = 8 };
enum { CI_TYPE = 0, CI_PHASE = 1, CI_ABS_START = 2, CI_ABS_N = 3, CI_EMS_START = 4, CI_EMS_N = 5, CI_GUIDE_START = 10,
       CI_HAS_GUIDE = 11 };
enum { RI_NODE = 0, RI_EVENT = 1, RI_HAS_FACET = 2, RI_HIST_START = 3, RI_HIST_N = 4 };
enum { HI_PROP_A = 0, HI_PROP_B = 1, HI_NA = 2, HI_NB = 3, HI_OFFSET = 4 };
enum { LI_POS = 0, LI_DIR = 1, LI_WL = 2, LI_WL_START = 3, LI_WL_N = 4 };
#endif

}  // namespace pvt
