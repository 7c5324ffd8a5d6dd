// pvt_api.cu -- the C ABI of include/pvtrace_b200.h on top of the kernels in pvt_kernels.cuh.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC  (csrc/build.py)
// There is no CPU path in this library: every compute entry fails with an error string if CUDA cannot run it.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pvtrace_b200.h"
#include "pvt_common.cuh"
#include "pvt_aux_kernels.cuh"
#include "pvt_launch.h"

namespace pvt {

static thread_local char g_error[1024] = "";
char* error_buffer() { return g_error; }
int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return 1;
}

}  // namespace pvt

using namespace pvt;

// ---------------------------------------------------------------------------------------------------------
// Context: the device-resident scene + tally accumulators + event-log buffers of ONE device.

struct pvt_context {
  int device = 0;
  int sm_count = 0;
  Header hdr;
  int blob_words = 0;
  int scene_in_smem = 1;
  size_t smem_bytes = 0;
  int has_emitter = 0;
  std::vector<double> host_blob;
  DeviceBuffer<double> blob;
  // one allocation of 8-byte words: [distinct R | cross R | sums 8R | bins B | stats NSTATS | work counter 1]
  DeviceBuffer<u64> tallies;
  size_t tally_words = 0;
  DeviceBuffer<double> packed;
  DeviceBuffer<uint32_t> arrived;  // streaming-upload mark (see TraceArgs::arrived)
  DeviceBuffer<u64> slabs;  // CTA-private tally slabs of one launch: [max_grid][10 R]
  DeviceBuffer<double> requests;  // tally-request rings of the service-warp kernels: [resident CTAs][kReqWords][pool]
  int wave_service = 0;     // service threads per CTA (0: tallies are made in place)
  int max_grid = 0;
  int wave_threads = 0;     // CTA size of the wavefront kernel for this scene, 0: scene needs trace_kernel
  int wave_pool = 0;        // photon slots per CTA
  int wave_ctas = 1;        // resident CTAs per SM
  bool wave_boxes = false;  // every node is an axis-aligned box: the kBoxes instantiation (no primitive switch)
  size_t wave_smem = 0;
  // event log of the last trace
  long long log_rows = 0, log_rays = 0;
  DeviceBuffer<int32_t> counts, hit, container, adjacent, component, source;
  DeviceBuffer<uint8_t> kind;
  DeviceBuffer<double> position, direction, normal, wavelength, travelled, duration;
  int blocks_per_sm[4] = {0, 0, 0, 0};  // per kernel instantiation
  long long launches = 0;

  int R() const { return hdr.n_recorders; }
  int B() const { return hdr.total_bins; }
  u64* d_distinct() { return tallies.ptr; }
  u64* d_cross() { return tallies.ptr + R(); }
  double* d_sums() { return reinterpret_cast<double*>(tallies.ptr + 2 * R()); }
  u64* d_bins() { return tallies.ptr + 10 * R(); }
  u64* d_stats() { return tallies.ptr + 10 * R() + B(); }
  u64* d_work() { return tallies.ptr + 10 * R() + B() + PVT_NSTATS; }
};

static int validate_scene(const pvt_scene_t* S) {
  if (!S) return fail("scene is NULL");
  if (S->n_nodes <= 0) return fail("scene has no geometry nodes");
  if (S->n_nodes > PVT_MAX_NODES) return fail("Engine supports at most %d geometry nodes.", PVT_MAX_NODES);
  if (S->n_recorders > PVT_MAX_RECORDERS) return fail("at most %d recorders are supported", PVT_MAX_RECORDERS);
  if (S->root_id < 0 || S->root_id >= S->n_nodes) return fail("root_id out of range");
  for (int i = 0; i < S->n_nodes; ++i) {
    if (S->geom_type[i] < 0 || S->geom_type[i] > 2) return fail("node %d: unknown geometry tag %d", i, S->geom_type[i]);
    if (S->comp_start[i] < 0 || S->comp_start[i] + S->comp_count[i] > S->n_components)
      return fail("node %d: component range out of bounds", i);
  }
  for (int c = 0; c < S->n_components; ++c) {
    if (S->comp_abs_n[c] < 1 || S->comp_abs_start[c] < 0 || S->comp_abs_start[c] + S->comp_abs_n[c] > S->n_abs_knots)
      return fail("component %d: absorption table out of bounds", c);
    if (S->comp_type[c] == PVT_COMP_LUMINOPHORE &&
        (S->comp_ems_n[c] < 1 || S->comp_ems_start[c] < 0 || S->comp_ems_start[c] + S->comp_ems_n[c] > S->n_ems_knots))
      return fail("component %d: emission table out of bounds", c);
  }
  for (int r = 0; r < S->n_recorders; ++r) {
    if (S->rec_node[r] < 0 || S->rec_node[r] >= S->n_nodes) return fail("recorder %d: node out of range", r);
    if (S->rec_hist_start[r] < 0 || S->rec_hist_start[r] + S->rec_hist_n[r] > S->n_hists)
      return fail("recorder %d: histogram range out of bounds", r);
  }
  if (S->n_facets > 0 && S->facet_count) {
    for (int i = 0; i < S->n_nodes; ++i)
      if (S->facet_start[i] < 0 || S->facet_count[i] < 0 || S->facet_start[i] + S->facet_count[i] > S->n_facets)
        return fail("node %d: facet range out of bounds", i);
    if (S->facet_refl_n && S->refl_x && S->refl_y)
      for (int f = 0; f < S->n_facets; ++f)
        if (S->facet_refl_n[f] < 0 || (S->facet_refl_n[f] > 0 && (S->facet_refl_start[f] < 0 ||
                                                                 S->facet_refl_start[f] + S->facet_refl_n[f] > S->n_refl_knots)))
          return fail("facet %d: reflectivity table out of bounds", f);
  }
  for (int h = 0; h < S->n_hists; ++h) {
    const long long cells = (long long)S->hist_na[h] * (S->hist_prop_b[h] < 0 ? 1 : S->hist_nb[h]);
    if (S->hist_offset[h] < 0 || S->hist_offset[h] + cells > S->total_bins) return fail("histogram %d: bins out of bounds", h);
  }
  return 0;
}

extern "C" int pvt_context_create(const pvt_scene_t* scene, const pvt_emit_t* emit, int device, pvt_context_t** out) {
  if (!out) return fail("ctx out pointer is NULL");
  *out = nullptr;
  PVT_TRY(validate_scene(scene));
  if (emit && emit->n_lights <= 0) return fail("emitter has no lights");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return fail("no CUDA device is usable (pvtrace_b200 has no CPU fallback)");
  if (device < 0 || device >= n_dev) return fail("device %d out of range (have %d)", device, n_dev);
  PVT_CUDA(cudaSetDevice(device));

  pvt_context* c = new pvt_context();
  c->device = device;
  c->host_blob = pack_scene(*scene, emit);
  memcpy(&c->hdr, c->host_blob.data(), sizeof(Header));
  c->blob_words = c->hdr.total_words;
  c->has_emitter = emit != nullptr;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { delete c; return fail("cudaGetDeviceProperties -> %s", cudaGetErrorString(e)); }
  c->sm_count = prop.multiProcessorCount;
  // the blob lives in shared memory when it leaves room for >= 2 CTAs per SM; otherwise it is read through L1
  const size_t want = trace_smem_bytes(c->blob_words);
  c->scene_in_smem = want <= (size_t)prop.sharedMemPerBlockOptin && want <= 100 * 1024;
  c->smem_bytes = trace_smem_bytes(c->scene_in_smem ? c->blob_words : 0);
  // wavefront kernel: needs the blob AND the photon pool in shared memory, <= 64 recorders (seen mask), <= 254 nodes
  c->wave_threads = 0;
  c->wave_boxes = !getenv("PVT_NO_BOX_KERNEL");
  for (int i = 0; i < scene->n_nodes; ++i) {
    const double* m = scene->world_to_local + 16 * i;
    const bool aligned = m[0] == 1.0 && m[1] == 0.0 && m[2] == 0.0 && m[4] == 0.0 && m[5] == 1.0 && m[6] == 0.0 &&
                         m[8] == 0.0 && m[9] == 0.0 && m[10] == 1.0;
    if (scene->geom_type[i] != PVT_GEOM_BOX || !aligned) c->wave_boxes = false;
  }
  if (c->R() <= 64) {
    int want_t = 0, want_p = 0, want_b = 0;  // 0: any
    if (const char* env = getenv("PVT_WAVEFRONT_THREADS")) want_t = atoi(env);
    if (const char* env = getenv("PVT_WAVEFRONT_POOL")) want_p = atoi(env);
    if (const char* env = getenv("PVT_WAVEFRONT_CTAS")) want_b = atoi(env);
    for (int k = 0; k < wave_variant_count() && want_t >= 0; ++k) {
      const int t = wave_variant(k).threads, pl = wave_variant(k).pool, b = wave_variant(k).ctas;
      if ((want_t && t != want_t) || (want_p && pl != want_p) || (want_b && b != want_b)) continue;
      size_t need = wavefront_smem_bytes(c->blob_words, pl, t);
      if (const char* env = getenv("PVT_EXTRA_SMEM")) need += (size_t)atoi(env);  // experiment: shrink L1
      if ((need + 1024) * b <= (size_t)prop.sharedMemPerMultiprocessor && need <= (size_t)prop.sharedMemPerBlockOptin) {
        c->wave_threads = t; c->wave_pool = pl; c->wave_ctas = b; c->wave_smem = need;
        break;
      }
    }
  }

  int rc = c->blob.reserve((size_t)c->blob_words);
  c->tally_words = (size_t)10 * c->R() + c->B() + PVT_NSTATS + 1;
  if (!rc) rc = c->tallies.reserve(c->tally_words);
  if (!rc) rc = c->packed.reserve((size_t)10 * c->R() + c->B() + 1);
  c->max_grid = c->sm_count * 8;
  if (!rc) rc = c->slabs.reserve((size_t)c->max_grid * 10 * c->R() + 1);
  if (!rc) rc = c->arrived.reserve(4);
  c->wave_service = (c->wave_threads == 512 && c->wave_pool == 1024 && c->wave_ctas == 1 && c->R() > 0 &&
                     !(getenv("PVT_TALLY_IN_PLACE") && atoi(getenv("PVT_TALLY_IN_PLACE")))) ? wave_service_threads() : 0;
  if (!rc && c->wave_service) rc = c->requests.reserve((size_t)c->sm_count * c->wave_ctas * 2 * kReqWords * c->wave_pool);
  if (!rc && cudaMemcpy(c->blob.ptr, c->host_blob.data(), (size_t)c->blob_words * 8, cudaMemcpyHostToDevice) != cudaSuccess)
    rc = fail("scene upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (!rc && cudaMemset(c->tallies.ptr, 0, c->tally_words * 8) != cudaSuccess)
    rc = fail("tally reset failed: %s", cudaGetErrorString(cudaGetLastError()));
  for (int which = 0; which < 4 && !rc; ++which) rc = reg_occupancy(which, c->smem_bytes, &c->blocks_per_sm[which]);
  if (!rc && c->wave_threads > 0) rc = wave_setup(WaveVariant{c->wave_threads, c->wave_pool, c->wave_ctas}, c->wave_smem);
  if (rc) {
    pvt_context_destroy(c);
    return rc;
  }
  *out = c;
  return 0;
}

extern "C" int pvt_context_destroy(pvt_context_t* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  delete c;  // every DeviceBuffer member frees itself
  return 0;
}

extern "C" int pvt_context_reset(pvt_context_t* c, void* stream) {
  if (!c) return fail("ctx is NULL");
  PVT_CUDA(cudaSetDevice(c->device));
  PVT_CUDA(cudaMemsetAsync(c->tallies.ptr, 0, c->tally_words * 8, (cudaStream_t)stream));
  c->launches = 0;
  return 0;
}

static int prepare_log(pvt_context* c, const pvt_params_t* P, cudaStream_t st) {
  c->log_rows = 0;
  c->log_rays = 0;
  if (P->record_every <= 0) return 0;
  const long long rays = (P->n + P->record_every - 1) / P->record_every;
  const long long rows = rays * (long long)P->max_events;
  PVT_TRY(c->counts.reserve((size_t)rays));
  PVT_TRY(c->kind.reserve((size_t)rows));
  PVT_TRY(c->hit.reserve((size_t)rows));
  PVT_TRY(c->container.reserve((size_t)rows));
  PVT_TRY(c->adjacent.reserve((size_t)rows));
  PVT_TRY(c->component.reserve((size_t)rows));
  PVT_TRY(c->source.reserve((size_t)rows));
  PVT_TRY(c->position.reserve((size_t)rows * 3));
  PVT_TRY(c->direction.reserve((size_t)rows * 3));
  PVT_TRY(c->normal.reserve((size_t)rows * 3));
  PVT_TRY(c->wavelength.reserve((size_t)rows));
  PVT_TRY(c->travelled.reserve((size_t)rows));
  PVT_TRY(c->duration.reserve((size_t)rows));
  // initial values of the reference's log arrays (_kernel.pyx:1035-1047): ids -1, everything else 0
  PVT_CUDA(cudaMemsetAsync(c->counts.ptr, 0, (size_t)rays * 4, st));
  PVT_CUDA(cudaMemsetAsync(c->kind.ptr, 0, (size_t)rows, st));
  PVT_CUDA(cudaMemsetAsync(c->hit.ptr, 0xFF, (size_t)rows * 4, st));
  PVT_CUDA(cudaMemsetAsync(c->container.ptr, 0xFF, (size_t)rows * 4, st));
  PVT_CUDA(cudaMemsetAsync(c->adjacent.ptr, 0xFF, (size_t)rows * 4, st));
  PVT_CUDA(cudaMemsetAsync(c->component.ptr, 0xFF, (size_t)rows * 4, st));
  PVT_CUDA(cudaMemsetAsync(c->source.ptr, 0xFF, (size_t)rows * 4, st));
  PVT_CUDA(cudaMemsetAsync(c->position.ptr, 0, (size_t)rows * 24, st));
  PVT_CUDA(cudaMemsetAsync(c->direction.ptr, 0, (size_t)rows * 24, st));
  PVT_CUDA(cudaMemsetAsync(c->normal.ptr, 0, (size_t)rows * 24, st));
  PVT_CUDA(cudaMemsetAsync(c->wavelength.ptr, 0, (size_t)rows * 8, st));
  PVT_CUDA(cudaMemsetAsync(c->travelled.ptr, 0, (size_t)rows * 8, st));
  PVT_CUDA(cudaMemsetAsync(c->duration.ptr, 0, (size_t)rows * 8, st));
  c->log_rows = rows;
  c->log_rays = rays;
  return 0;
}

static int check_params(const pvt_params_t* P) {
  if (!P) return fail("params is NULL");
  if (P->n < 0) return fail("n must be >= 0");
  if (P->emit_method < 0 || P->emit_method > 2) return fail("emit_method must be 0 (kT), 1 (redshift) or 2 (full)");
  if (P->rng_mode != PVT_RNG_PHILOX && P->rng_mode != PVT_RNG_XOSHIRO) return fail("unknown rng_mode %d", P->rng_mode);
  if (P->record_every < 0) return fail("record_every must be >= 0");
  if (P->record_every > 0 && P->max_events < 2) return fail("max_events must be >= 2 when rays are recorded");
  return 0;
}

// CTAs of the wavefront kernel for a bundle of n rays, 0 when the bundle has to go through trace_kernel
static int wave_grid(const pvt_context* c, const pvt_params_t* P) {
  if (c->wave_threads <= 0 || P->rng_mode != PVT_RNG_PHILOX || (P->flags & PVT_FLAG_REGISTER_KERNEL) || P->n <= 0) return 0;
  if (P->n >= (1ll << 32) - 1024) return 0;  // the pool addresses photons of a bundle with 32 bits
  const long long want_blocks = (P->n + c->wave_pool - 1) / c->wave_pool;
  const long long resident = (long long)c->sm_count * c->wave_ctas;
  return (int)(want_blocks < resident ? want_blocks : resident);
}

// Columns of a host bundle that hold one value for every ray (see "constant columns" below): the value goes to the
// kernel in TraceArgs instead of the column going over PCIe.
namespace {
struct RayConstants {
  uint32_t mask = 0;  // bit 0: positions, 1: directions, 2: wavelengths
  double pos[3] = {0, 0, 0}, dir[3] = {0, 0, 0}, wl = 0;
};
}  // namespace

static int trace_device_impl(pvt_context_t* c, const double* d_pos, const double* d_dir, const double* d_wl,
                             const pvt_params_t* P, void* stream, const uint32_t* arrived, const RayConstants* consts);

extern "C" int pvt_trace_device(pvt_context_t* c, const double* d_pos, const double* d_dir, const double* d_wl,
                                const pvt_params_t* P, void* stream) {
  return trace_device_impl(c, d_pos, d_dir, d_wl, P, stream, nullptr, nullptr);
}

static int trace_device_impl(pvt_context_t* c, const double* d_pos, const double* d_dir, const double* d_wl,
                             const pvt_params_t* P, void* stream, const uint32_t* arrived, const RayConstants* consts) {
  if (!c) return fail("ctx is NULL");
  PVT_TRY(check_params(P));
  const uint32_t cmask = consts ? consts->mask : 0u;
  const bool have_rays = (d_pos || (cmask & 1u)) && (d_dir || (cmask & 2u)) && (d_wl || (cmask & 4u));
  if (!have_rays && !c->has_emitter) return fail("no ray arrays given and the context has no emitter");
  PVT_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  PVT_TRY(prepare_log(c, P, st));
  if (P->n == 0) return 0;
  PVT_CUDA(cudaMemsetAsync(c->d_work(), 0, 8, st));

  TraceArgs a;
  a.hdr = c->hdr;
  a.blob = c->blob.ptr; a.blob_words = c->blob_words; a.scene_in_smem = c->scene_in_smem;
  a.pos = have_rays && !(cmask & 1u) ? d_pos : nullptr;
  a.dir = have_rays && !(cmask & 2u) ? d_dir : nullptr;
  a.wl = have_rays && !(cmask & 4u) ? d_wl : nullptr;
  a.const_mask = have_rays ? cmask : 0u;
  for (int k = 0; k < 3; ++k) { a.cpos[k] = consts ? consts->pos[k] : 0.0; a.cdir[k] = consts ? consts->dir[k] : 0.0; }
  a.cwl = consts ? consts->wl : 0.0;
  a.n = P->n; a.first_index = P->first_index; a.record_every = P->record_every; a.keys = make_run_seed(P->seed);
  a.sp.maxsteps = P->maxsteps; a.sp.max_events = P->max_events; a.sp.emit_method = P->emit_method;
  a.work_counter = c->d_work();
  a.g_distinct = c->d_distinct(); a.g_cross = c->d_cross(); a.g_sums = c->d_sums(); a.g_bins = c->d_bins();
  a.g_stats = c->d_stats();
  a.log = LogColumns{c->counts.ptr, c->kind.ptr, c->hit.ptr, c->container.ptr, c->adjacent.ptr, c->component.ptr,
                     c->source.ptr, c->position.ptr, c->direction.ptr, c->normal.ptr, c->wavelength.ptr,
                     c->travelled.ptr, c->duration.ptr};

  a.slabs = c->slabs.ptr;
  a.requests = c->requests.ptr;
  int grid = wave_grid(c, P);
  a.arrived = arrived;
  if (grid > 0) {
    // persistent CTAs that claim blocks of photons from the work counter (zeroed above) as their pools drain
    PVT_CUDA(cudaMemsetAsync(c->slabs.ptr, 0, (size_t)grid * 10 * c->R() * 8 + 8, st));
    PVT_TRY(wave_launch(WaveVariant{c->wave_threads, c->wave_pool, c->wave_ctas}, c->wave_service, c->wave_boxes,
                        P->record_every > 0, a, grid, c->wave_smem, st));
  } else {
    if (arrived) return fail("streaming upload needs the wavefront kernel");
    const int wide = c->R() > 64;
    const int which = (P->rng_mode == PVT_RNG_XOSHIRO ? 2 : 0) + wide;
    long long want_blocks = (P->n + kTraceThreads - 1) / kTraceThreads;
    long long resident = (long long)c->sm_count * c->blocks_per_sm[which];
    if (resident > c->max_grid) resident = c->max_grid;
    grid = (int)(want_blocks < resident ? want_blocks : resident);
    PVT_CUDA(cudaMemsetAsync(c->slabs.ptr, 0, (size_t)grid * 10 * c->R() * 8 + 8, st));
    PVT_TRY(reg_launch(which, a, grid, c->smem_bytes, st));
  }
  PVT_CUDA(cudaGetLastError());
  c->launches += 1;
  return 0;
}

template <class T>
static int fetch(T* host, const T* dev, size_t count, cudaStream_t st) {
  if (!host || count == 0) return 0;
  PVT_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
  return 0;
}

extern "C" int pvt_context_read(pvt_context_t* c, pvt_out_t* out, void* stream) {
  if (!c || !out) return fail("ctx/out is NULL");
  PVT_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int R = c->R(), B = c->B();
  // int64 <- u64 is a reinterpretation; counts never approach 2^63
  PVT_TRY(fetch(reinterpret_cast<u64*>(out->rec_distinct), c->d_distinct(), (size_t)R, st));
  PVT_TRY(fetch(reinterpret_cast<u64*>(out->rec_crossings), c->d_cross(), (size_t)R, st));
  PVT_TRY(fetch(out->rec_sums, c->d_sums(), (size_t)8 * R, st));
  PVT_TRY(fetch(reinterpret_cast<u64*>(out->rec_bins), c->d_bins(), (size_t)B, st));
  if (out->stats) PVT_TRY(fetch(reinterpret_cast<u64*>(out->stats), c->d_stats(), (size_t)PVT_NSTATS, st));
  if (c->log_rows > 0 && out->kind) {
    const size_t rows = (size_t)c->log_rows;
    PVT_TRY(fetch(out->counts, c->counts.ptr, (size_t)c->log_rays, st));
    PVT_TRY(fetch(out->kind, c->kind.ptr, rows, st));
    PVT_TRY(fetch(out->hit, c->hit.ptr, rows, st));
    PVT_TRY(fetch(out->container, c->container.ptr, rows, st));
    PVT_TRY(fetch(out->adjacent, c->adjacent.ptr, rows, st));
    PVT_TRY(fetch(out->component, c->component.ptr, rows, st));
    PVT_TRY(fetch(out->source, c->source.ptr, rows, st));
    PVT_TRY(fetch(out->position, c->position.ptr, rows * 3, st));
    PVT_TRY(fetch(out->direction, c->direction.ptr, rows * 3, st));
    PVT_TRY(fetch(out->normal, c->normal.ptr, rows * 3, st));
    PVT_TRY(fetch(out->wavelength, c->wavelength.ptr, rows, st));
    PVT_TRY(fetch(out->travelled, c->travelled.ptr, rows, st));
    PVT_TRY(fetch(out->duration, c->duration.ptr, rows, st));
  }
  PVT_CUDA(cudaStreamSynchronize(st));
  if (out->stats) out->stats[PVT_STAT_LAUNCHES] = c->launches;
  return 0;
}

extern "C" int pvt_context_pack_tallies(pvt_context_t* c, double** d_packed, int64_t* n_packed, void* stream) {
  if (!c || !d_packed || !n_packed) return fail("NULL argument");
  PVT_CUDA(cudaSetDevice(c->device));
  const int R = c->R(), B = c->B();
  const int total = 10 * R + B;
  if (total > 0) {
    pack_tallies_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c->d_distinct(), 2 * R, c->d_sums(), 8 * R,
                                                                               c->d_bins(), B, c->packed.ptr);
    PVT_CUDA(cudaGetLastError());
  }
  *d_packed = c->packed.ptr;
  *n_packed = total;
  return 0;
}

extern "C" int pvt_context_unpack_tallies(pvt_context_t* c, void* stream) {
  if (!c) return fail("ctx is NULL");
  PVT_CUDA(cudaSetDevice(c->device));
  const int R = c->R(), B = c->B();
  const int total = 10 * R + B;
  if (total > 0) {
    unpack_tallies_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c->d_distinct(), 2 * R, c->d_sums(), 8 * R,
                                                                                 c->d_bins(), B, c->packed.ptr);
    PVT_CUDA(cudaGetLastError());
  }
  return 0;
}

static int grid_for(long long n, int sm_count) {
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

extern "C" int pvt_emit_device(pvt_context_t* c, double* d_pos, double* d_dir, double* d_wl, int64_t n, int64_t first_index,
                               uint64_t seed, void* stream) {
  if (!c) return fail("ctx is NULL");
  if (!c->has_emitter) return fail("the context has no emitter");
  if (n <= 0) return 0;
  PVT_CUDA(cudaSetDevice(c->device));
  emit_kernel<<<grid_for(n, c->sm_count), 256, 0, (cudaStream_t)stream>>>(c->blob.ptr, d_pos, d_dir, d_wl, n, first_index, make_run_seed(seed));
  PVT_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pvt_intersect_device(pvt_context_t* c, const double* d_pos, const double* d_dir, int64_t n, double* d_t0,
                                    int32_t* d_hit, int32_t* d_container, int32_t* d_adjacent, void* stream) {
  if (!c) return fail("ctx is NULL");
  if (n <= 0) return 0;
  PVT_CUDA(cudaSetDevice(c->device));
  return intersect_launch(c->wave_boxes, false, c->hdr, c->blob.ptr, d_pos, d_dir, n, d_t0, nullptr, d_hit, d_container,
                          d_adjacent, c->sm_count, (cudaStream_t)stream);
}

extern "C" int pvt_intersect_device_packed(pvt_context_t* c, const double* d_pos, const double* d_dir, int64_t n, double* d_t0,
                                           uint32_t* d_ids, void* stream) {
  if (!c) return fail("ctx is NULL");
  if (n <= 0) return 0;
  PVT_CUDA(cudaSetDevice(c->device));
  return intersect_launch(c->wave_boxes, true, c->hdr, c->blob.ptr, d_pos, d_dir, n, d_t0, d_ids, nullptr, nullptr, nullptr,
                          c->sm_count, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------
// Host-buffer entry points

namespace {
struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  int start() {
    PVT_CUDA(cudaEventCreate(&a));
    PVT_CUDA(cudaEventCreate(&b));
    PVT_CUDA(cudaEventRecord(a, 0));
    return 0;
  }
  int stop(double* seconds) {
    PVT_CUDA(cudaEventRecord(b, 0));
    PVT_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    PVT_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (seconds) *seconds = ms * 1e-3;
    return 0;
  }
  ~Timer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

constexpr int kMaxChunks = 64;           // pieces a host bundle is uploaded + traced in
constexpr size_t kMinChunkRays = 1 << 20;  // ... of at least this many rays each (separate launches)
constexpr size_t kStreamChunkRays = 1 << 17;  // smallest chunk of the streaming upload (one launch, arrival marks)
constexpr int kMaxDevices = 64;

// What the host-buffer entry points keep PER DEVICE: the cached context (bundles of the same scene -- engine.simulate_stream,
// repeated simulate calls -- reuse the uploaded blob and every device buffer), the ray staging buffer, two non-blocking
// streams and the upload events.  One bundle at a time per device (the mutex); different devices run concurrently, which
// is how pvt_trace_bundle_devices drives several GPUs from one process.
struct HostPath {
  std::mutex mutex;
  pvt_context* ctx = nullptr;
  DeviceBuffer<double> rays;  // staging for host rays: [pos 3n | dir 3n | wl n]
  cudaStream_t s_copy = nullptr, s_run = nullptr;
  cudaEvent_t uploaded[kMaxChunks] = {};
  uint32_t* h_marks = nullptr;  // page-locked: the arrival marks must not be staged
  // The streaming upload needs the copy stream to make progress WHILE the trace kernel runs.  Anything that serialises
  // the two (CUDA_LAUNCH_BLOCKING, a profiler replaying kernels one at a time) would leave the kernel polling until it
  // gives up, so the path is switched off when such a tool is detected, when the user says so (PVT_STREAM_UPLOAD=0), and
  // for good after a bundle that did not complete (which is then re-traced the plain way).
  bool stream_upload_ok = true;
};
HostPath g_paths[kMaxDevices];

bool stream_upload_allowed(const HostPath& path) {
  if (!path.stream_upload_ok) return false;
  if (const char* env = getenv("PVT_STREAM_UPLOAD")) return atoi(env) != 0;
  if (const char* env = getenv("CUDA_LAUNCH_BLOCKING")) { if (atoi(env) != 0) return false; }
  if (getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NVTX_INJECTION64_PATH"))
    return false;
  return true;
}

int acquire_path(int device, HostPath** out) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return fail("no CUDA device is usable (pvtrace_b200 has no CPU fallback)");
  if (device < 0 || device >= n_dev || device >= kMaxDevices) return fail("device %d out of range (have %d)", device, n_dev);
  *out = &g_paths[device];
  return 0;
}

// (call with path.mutex held)
int acquire_context(HostPath& path, const pvt_scene_t* scene, const pvt_emit_t* emit, int device, pvt_context** out) {
  PVT_TRY(validate_scene(scene));
  std::vector<double> blob = pack_scene(*scene, emit);
  pvt_context*& cached = path.ctx;
  if (cached && cached->host_blob.size() == blob.size() &&
      memcmp(cached->host_blob.data(), blob.data(), blob.size() * 8) == 0 && cached->has_emitter == (emit != nullptr)) {
    *out = cached;
    return 0;
  }
  if (cached) { pvt_context_destroy(cached); cached = nullptr; }
  PVT_TRY(pvt_context_create(scene, emit, device, &cached));
  *out = cached;
  return 0;
}

int ensure_streams(HostPath& path) {
  if (path.s_copy) return 0;
  PVT_CUDA(cudaStreamCreateWithFlags(&path.s_copy, cudaStreamNonBlocking));
  PVT_CUDA(cudaStreamCreateWithFlags(&path.s_run, cudaStreamNonBlocking));
  for (auto& e : path.uploaded) PVT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  PVT_CUDA(cudaHostAlloc((void**)&path.h_marks, kMaxChunks * sizeof(uint32_t), cudaHostAllocDefault));
  return 0;
}

// ---- constant columns --------------------------------------------------------------------------------------
// A bundle from a point source carries the same position in every row, a monochromatic one the same wavelength, a
// collimated one the same direction: 24 + 8 (+ 24) of the 56 bytes per ray that cross PCIe say nothing.  The host
// finds such columns and passes ONE value to the kernel instead (TraceArgs::const_mask).  Finding them means reading
// every row, and at host-memory speed that is as long as the upload it saves -- so it is done OPTIMISTICALLY: a column
// whose first rows agree is taken as constant, the trace starts at once on the remaining columns, and worker threads
// check the rest of the column while the device is busy.  In the (contrived) case that a later row differs, the result
// is discarded and the bundle traced again with every column uploaded.
constexpr size_t kProbeRows = 4096;       // rows compared up front
constexpr size_t kElideMinRays = 1 << 18;  // smaller bundles are not worth the threads
constexpr size_t kScanBlockRows = 1 << 14;

class ConstantScan {
 public:
  RayConstants found;
  // decides the candidates from the first rows (cheap: the upload can be planned at once) ...
  void probe(const double* pos, const double* dir, const double* wl, size_t n) {
    cols_[0] = Column{pos, 3}; cols_[1] = Column{dir, 3}; cols_[2] = Column{wl, 1};
    n_ = n;
    const size_t probe = n < kProbeRows ? n : kProbeRows;
    for (int k = 0; k < 3; ++k) {
      fill_template(k);
      if (rows_equal(k, 0, probe)) found.mask |= 1u << k;
    }
    for (int k = 0; k < 3; ++k) { found.pos[k] = pos[k]; found.dir[k] = dir[k]; }
    found.wl = wl[0];
    blocks_per_col_ = (n + kScanBlockRows - 1) / kScanBlockRows;
    next_.store(0);
    failed_.store(false);
  }
  // ... and checks the rest of the candidate columns on `threads - 1` workers (started once the copies are enqueued:
  // creating threads takes as long as uploading a million rays)
  void start(int threads) {
    if (!found.mask) return;
    for (int t = 1; t < threads; ++t) workers_.emplace_back([this] { work(); });
  }
  // the calling thread helps, then joins the workers; false: some candidate column was not constant after all
  bool finish() {
    if (found.mask) work();
    for (auto& w : workers_) w.join();
    workers_.clear();
    return !failed_.load();
  }
  ~ConstantScan() { for (auto& w : workers_) if (w.joinable()) w.join(); }

 private:
  struct Column { const double* data; int width; };
  static constexpr size_t kTemplateRows = 512;
  Column cols_[3];
  double tmpl_[3][kTemplateRows * 3];
  size_t n_ = 0, blocks_per_col_ = 0;
  std::atomic<size_t> next_{0};
  std::atomic<bool> failed_{false};
  std::vector<std::thread> workers_;

  void fill_template(int k) {
    const int w = cols_[k].width;
    for (size_t r = 0; r < kTemplateRows; ++r)
      for (int j = 0; j < w; ++j) tmpl_[k][r * w + j] = cols_[k].data[j];
  }
  // bitwise comparison (memcmp is vectorised): -0.0 differs from 0.0 and a NaN equals itself, which is what is wanted
  bool rows_equal(int k, size_t lo, size_t hi) const {
    const int w = cols_[k].width;
    for (size_t r = lo; r < hi; r += kTemplateRows) {
      const size_t m = hi - r < kTemplateRows ? hi - r : kTemplateRows;
      if (memcmp(cols_[k].data + r * w, tmpl_[k], m * w * sizeof(double)) != 0) return false;
    }
    return true;
  }
  void work() {
    for (;;) {
      if (failed_.load(std::memory_order_relaxed)) return;
      const size_t b = next_.fetch_add(1);
      if (b >= 3 * blocks_per_col_) return;
      const int k = (int)(b / blocks_per_col_);
      if (!(found.mask & (1u << k))) continue;
      const size_t lo = (b % blocks_per_col_) * kScanBlockRows, hi = lo + kScanBlockRows < n_ ? lo + kScanBlockRows : n_;
      if (!rows_equal(k, lo, hi)) { failed_.store(true); return; }
    }
  }
};

int scan_threads() {
  if (const char* env = getenv("PVT_SCAN_THREADS")) { const int v = atoi(env); if (v >= 1 && v <= 64) return v; }
  // half the cores, at most 8 (four already keep up with one GPU's upload); ranks that share the host (torchrun sets
  // LOCAL_WORLD_SIZE) share its cores
  unsigned hw = std::thread::hardware_concurrency();
  if (const char* env = getenv("LOCAL_WORLD_SIZE")) {
    const int ranks = atoi(env);
    if (ranks > 1) hw = 2 * hw / (unsigned)ranks;
  }
  const unsigned want = hw / 2 < 8 ? hw / 2 : 8;
  return want < 2 ? 2 : (int)want;
}
bool elision_allowed() {
  if (const char* env = getenv("PVT_ELIDE_CONSTANT")) return atoi(env) != 0;
  return true;
}
}  // namespace

// One bundle on one device, host buffers in and out.  h2d_bytes: what actually crossed PCIe for the rays.
static int trace_host_bundle(const pvt_scene_t* scene, const pvt_emit_t* emit, const double* positions,
                             const double* directions, const double* wavelengths, const pvt_params_t* params,
                             pvt_out_t* out, double* elapsed_s) {
  PVT_TRY(check_params(params));
  if (!out) return fail("out is NULL");
  const bool have_rays = positions && directions && wavelengths;
  if (!have_rays && !emit) return fail("either ray arrays or an emitter are required");
  HostPath* path_ptr = nullptr;
  PVT_TRY(acquire_path(params->device, &path_ptr));
  HostPath& path = *path_ptr;
  std::lock_guard<std::mutex> lock(path.mutex);
  PVT_CUDA(cudaSetDevice(params->device));
  pvt_context* c = nullptr;
  PVT_TRY(acquire_context(path, scene, emit, params->device, &c));
  // two non-blocking streams: host rays are uploaded in chunks on one while the trace runs on the other (overlap needs
  // page-locked host memory; with pageable memory the copies simply serialise)
  PVT_TRY(ensure_streams(path));
  cudaStream_t s_copy = path.s_copy, s_run = path.s_run;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  PVT_CUDA(cudaEventCreate(&t0));
  PVT_CUDA(cudaEventCreate(&t1));
  PVT_CUDA(cudaEventRecord(t0, s_run));
  int rc = pvt_context_reset(c, s_run);
  const size_t n = (size_t)params->n;
  long long h2d_bytes = 0;
  // Opt-in (PVT_ZERO_COPY=1): page-locked host arrays (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) are
  // read by the kernel IN PLACE, the warps that fill the shared-memory ray ring pulling them over PCIe a whole ring
  // ahead of their use.  Measured 42-45 GB/s against the copy engine's 55 GB/s, so the streaming upload below
  // is the default.
  bool streamed = false;
  ConstantScan scan;
  const double *z_pos = nullptr, *z_dir = nullptr, *z_wl = nullptr;
  if (!rc && have_rays && n > 0 && getenv("PVT_ZERO_COPY") && atoi(getenv("PVT_ZERO_COPY")) == 1) {
    const void* host[3] = {positions, directions, wavelengths};
    const double* dev[3] = {nullptr, nullptr, nullptr};
    bool all = true;
    for (int k = 0; k < 3 && all; ++k) {
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, host[k]) != cudaSuccess) { cudaGetLastError(); all = false; break; }
      all = attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
      dev[k] = static_cast<const double*>(attr.devicePointer);
    }
    if (all) { z_pos = dev[0]; z_dir = dev[1]; z_wl = dev[2]; }
  }
  if (!rc && z_pos) {
    rc = pvt_trace_device(c, z_pos, z_dir, z_wl, params, s_run);
    h2d_bytes = 56ll * (long long)n;
  } else if (!rc && have_rays && n > 0 && wave_grid(c, params) > 0 && stream_upload_allowed(path)) {
    // Streaming upload: the trace kernel starts at once and polls an arrival mark; the copy engine delivers the
    // arrays front to back in chunks (plain copies of the columns that are not constant) followed, in stream order, by
    // the new mark = rays complete so far.  CTAs claim photons in index order, so they consume the prefix as it lands.
    // Upload and trace overlap completely: total time ~ max(PCIe, kernel) + the trace of the LAST chunk -- hence chunks
    // that start small (the kernel gets going at once), grow up to a fifth of what is left (few, efficient copies) and
    // shrink again towards the end (little left to trace when the last byte lands).  They grow by a tenth, not by
    // doubling: the kernel traces chunk k while chunk k + 1 lands, so a chunk must not take longer to upload than its
    // predecessor takes to trace.  With the constant columns gone the upload is only ~1.15x faster than the trace
    // (240 MB in 4.4 ms against 5.0 ms), and doubling chunks left the kernel waiting at every step of the ramp
    // (measured with the kernel at 6.2 ms: trace done at 6.9 ms instead of 6.3; with it at 5.0 ms, growth 105 / 110 /
    // 115 / 125 / 150 % from 64 K rays: 5.65 / 5.60 / 5.58 / 5.62 / 5.70 ms, 110 % from 128 K rays: 5.53 ms).
    streamed = true;
    if (n >= kElideMinRays && elision_allowed()) scan.probe(positions, directions, wavelengths, n);
    const RayConstants& consts = scan.found;
    rc = path.rays.reserve(7 * n);
    double *d_pos = path.rays.ptr, *d_dir = path.rays.ptr + 3 * n, *d_wl = path.rays.ptr + 6 * n;
    size_t min_chunk = kStreamChunkRays, shrink_div = 5, growth_pct = 110;
    if (const char* env = getenv("PVT_UPLOAD_GROWTH_PCT")) { const int v = atoi(env); if (v >= 100 && v <= 400) growth_pct = (size_t)v; }
    if (const char* env = getenv("PVT_UPLOAD_MIN_CHUNK")) { const long v = atol(env); if (v >= 1024) min_chunk = (size_t)v; }
    if (const char* env = getenv("PVT_UPLOAD_SHRINK")) { const int v = atoi(env); if (v >= 2 && v <= 64) shrink_div = (size_t)v; }
    int chunks = 0;
    cudaError_t e = cudaSuccess;
    uint32_t* h_marks = path.h_marks;
    const bool dbg_on = getenv("PVT_DEBUG_TIMING") != nullptr;  // how long the copy stream took, when the trace ended
    cudaEvent_t dbg[2] = {nullptr, nullptr};
    if (dbg_on) { cudaEventCreate(&dbg[0]); cudaEventCreate(&dbg[1]); cudaEventRecord(dbg[0], s_copy); }
    if (!rc) e = cudaMemsetAsync(c->arrived.ptr, 0, 4, s_copy);
    if (!rc && e == cudaSuccess) e = cudaEventRecord(path.uploaded[0], s_copy);
    if (!rc && e == cudaSuccess) e = cudaStreamWaitEvent(s_run, path.uploaded[0], 0);
    if (!rc && e == cudaSuccess) rc = trace_device_impl(c, d_pos, d_dir, d_wl, params, s_run, c->arrived.ptr, &consts);
    if (consts.mask == 7u) {  // nothing to upload at all: every ray "has arrived"
      h_marks[0] = (uint32_t)n;
      if (!rc && e == cudaSuccess) e = cudaMemcpyAsync(c->arrived.ptr, h_marks, 4, cudaMemcpyHostToDevice, s_copy);
    } else {
      size_t lo = 0, grow = min_chunk;
      while (lo < n && !rc && e == cudaSuccess) {
        const size_t left = n - lo;
        size_t m = left / shrink_div > min_chunk ? left / shrink_div : min_chunk;
        if (m > grow) m = grow;
        grow = grow * growth_pct / 100;
        if (m > left || chunks == kMaxChunks - 1 || left - m < min_chunk / 2) m = left;
        if (!(consts.mask & 1u)) {
          e = cudaMemcpyAsync(d_pos + 3 * lo, positions + 3 * lo, 24 * m, cudaMemcpyHostToDevice, s_copy);
          h2d_bytes += 24ll * (long long)m;
        }
        if (e == cudaSuccess && !(consts.mask & 2u)) {
          e = cudaMemcpyAsync(d_dir + 3 * lo, directions + 3 * lo, 24 * m, cudaMemcpyHostToDevice, s_copy);
          h2d_bytes += 24ll * (long long)m;
        }
        if (e == cudaSuccess && !(consts.mask & 4u)) {
          e = cudaMemcpyAsync(d_wl + lo, wavelengths + lo, 8 * m, cudaMemcpyHostToDevice, s_copy);
          h2d_bytes += 8ll * (long long)m;
        }
        lo += m;
        h_marks[chunks] = (uint32_t)lo;
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->arrived.ptr, h_marks + chunks, 4, cudaMemcpyHostToDevice, s_copy);
        ++chunks;
      }
    }
    scan.start(scan_threads());
    if (dbg_on) {
      cudaEventRecord(dbg[1], s_copy);
      cudaEventSynchronize(dbg[1]);
      const bool ok = scan.finish();
      cudaEvent_t done;
      cudaEventCreate(&done);
      cudaEventRecord(done, s_run);
      cudaEventSynchronize(done);
      float up = 0.f, t_start = 0.f, total = 0.f;
      cudaEventElapsedTime(&up, dbg[0], dbg[1]);
      cudaEventElapsedTime(&t_start, t0, dbg[0]);
      cudaEventElapsedTime(&total, t0, done);
      fprintf(stderr, "[pvt] device %d: constant columns mask %u (%s), %d upload chunks, %lld bytes; upload started %.3f ms after "
              "t0 and took %.3f ms; trace done at %.3f ms\n", params->device, consts.mask, ok ? "confirmed" : "REFUTED", chunks,
              h2d_bytes, t_start, up, total);
      cudaEventDestroy(dbg[0]); cudaEventDestroy(dbg[1]); cudaEventDestroy(done);
    }
    if (!rc && e != cudaSuccess) {
      // never leave the kernel polling: publish "everything arrived" so it drains, then report
      h_marks[kMaxChunks - 1] = 0xffffffffu;
      cudaMemcpyAsync(c->arrived.ptr, h_marks + kMaxChunks - 1, 4, cudaMemcpyHostToDevice, s_copy);
      rc = fail("ray upload failed: %s", cudaGetErrorString(e));
    }
  } else if (!rc && have_rays && n > 0) {
    rc = path.rays.reserve(7 * n);
    double *d_pos = path.rays.ptr, *d_dir = path.rays.ptr + 3 * n, *d_wl = path.rays.ptr + 6 * n;
    // the event log is indexed by the ray's position in the bundle, so logged bundles go in one piece
    int chunks = (params->record_every == 0 && n >= (size_t)kMinChunkRays * 2) ? (int)(n / kMinChunkRays) : 1;
    if (chunks > kMaxChunks) chunks = kMaxChunks;
    if (const char* env = getenv("PVT_UPLOAD_CHUNKS")) {  // tuning knob
      const int want = atoi(env);
      if (want >= 1 && want <= kMaxChunks && params->record_every == 0) chunks = want;
    }
    for (int k = 0; k < chunks && !rc; ++k) {
      const size_t lo = n * k / chunks, hi = n * (k + 1) / chunks, m = hi - lo;
      cudaError_t e = cudaMemcpyAsync(d_pos + 3 * lo, positions + 3 * lo, 24 * m, cudaMemcpyHostToDevice, s_copy);
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_dir + 3 * lo, directions + 3 * lo, 24 * m, cudaMemcpyHostToDevice, s_copy);
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_wl + lo, wavelengths + lo, 8 * m, cudaMemcpyHostToDevice, s_copy);
      if (e == cudaSuccess) e = cudaEventRecord(path.uploaded[k], s_copy);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s_run, path.uploaded[k], 0);
      if (e != cudaSuccess) { rc = fail("ray upload failed: %s", cudaGetErrorString(e)); break; }
      h2d_bytes += 56ll * (long long)m;
      pvt_params_t part = *params;
      part.n = (int64_t)m;
      part.first_index = params->first_index + (int64_t)lo;
      rc = pvt_trace_device(c, d_pos + 3 * lo, d_dir + 3 * lo, d_wl + lo, &part, s_run);
    }
  } else if (!rc) {
    rc = pvt_trace_device(c, nullptr, nullptr, nullptr, params, s_run);
  }
  // (the calling thread joins the check of the constant columns while the device works)
  const bool columns_ok = scan.finish();
  if (!rc) rc = pvt_context_read(c, out, s_run);
  if (!rc && n > 0) {  // every ray must have been traced
    u64 traced = 0;
    if (cudaMemcpy(&traced, c->d_stats() + PVT_STAT_RAYS, 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
      rc = fail("reading the ray counter failed: %s", cudaGetErrorString(cudaGetLastError()));
    } else if (streamed && (traced != (u64)n || !columns_ok)) {
      // Either the kernel gave up polling for rays (something serialised upload and trace: never stream again), or a
      // column taken as constant was not.  Trace again the plain way, every column uploaded.
      if (traced != (u64)n) path.stream_upload_ok = false;
      cudaStreamSynchronize(s_copy);
      double* r = path.rays.ptr;
      cudaError_t e = cudaMemcpyAsync(r, positions, 24 * n, cudaMemcpyHostToDevice, s_run);
      if (e == cudaSuccess) e = cudaMemcpyAsync(r + 3 * n, directions, 24 * n, cudaMemcpyHostToDevice, s_run);
      if (e == cudaSuccess) e = cudaMemcpyAsync(r + 6 * n, wavelengths, 8 * n, cudaMemcpyHostToDevice, s_run);
      if (e != cudaSuccess) rc = fail("ray upload failed: %s", cudaGetErrorString(e));
      h2d_bytes += 56ll * (long long)n;
      if (!rc) rc = pvt_context_reset(c, s_run);
      if (!rc) rc = pvt_trace_device(c, r, r + 3 * n, r + 6 * n, params, s_run);
      if (!rc) rc = pvt_context_read(c, out, s_run);
    } else if (traced != (u64)n) {
      rc = fail("traced %llu of %zu rays", traced, n);
    }
  }
  if (!rc && out->stats) out->stats[PVT_STAT_H2D_BYTES] = h2d_bytes;
  if (!rc) {
    cudaError_t e = cudaEventRecord(t1, s_run);
    if (e == cudaSuccess) e = cudaEventSynchronize(t1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, t0, t1);
    if (e != cudaSuccess) rc = fail("timing failed: %s", cudaGetErrorString(e));
    else if (elapsed_s) *elapsed_s = ms * 1e-3;
  } else {
    cudaStreamSynchronize(s_run);
    cudaStreamSynchronize(s_copy);
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  return rc;
}

extern "C" int pvt_trace_bundle(const pvt_scene_t* scene, const pvt_emit_t* emit, const double* positions,
                                const double* directions, const double* wavelengths, const pvt_params_t* params,
                                pvt_out_t* out, double* elapsed_s) {
  return trace_host_bundle(scene, emit, positions, directions, wavelengths, params, out, elapsed_s);
}

// Several devices, one process: contiguous index slices of the bundle, one host thread per device (each with its own
// cached context, streams and staging), tallies summed on the host -- they have to come to the host anyway, 62 KB per
// device for the LSC configs -- and every device writing the log rows of its own slice straight into the caller's arrays.
extern "C" int pvt_trace_bundle_devices(const pvt_scene_t* scene, const pvt_emit_t* emit, const double* positions,
                                        const double* directions, const double* wavelengths, const pvt_params_t* params,
                                        int32_t n_devices, const int32_t* device_ids, pvt_out_t* out, double* elapsed_s) {
  PVT_TRY(check_params(params));
  if (!out) return fail("out is NULL");
  if (n_devices <= 0 || !device_ids) return fail("n_devices must be >= 1 and device_ids non-NULL");
  if (n_devices > kMaxDevices) return fail("at most %d devices", kMaxDevices);
  for (int a = 0; a < n_devices; ++a)
    for (int b = a + 1; b < n_devices; ++b)
      if (device_ids[a] == device_ids[b]) return fail("device %d listed twice", device_ids[a]);
  if (n_devices == 1 || params->n == 0) {
    pvt_params_t one = *params;
    one.device = device_ids[0];
    return trace_host_bundle(scene, emit, positions, directions, wavelengths, &one, out, elapsed_s);
  }
  PVT_TRY(validate_scene(scene));
  const bool have_rays = positions && directions && wavelengths;
  const int R = scene->n_recorders, B = scene->total_bins;
  const long long n = params->n, every = params->record_every;
  // slice boundaries: equal shares, rounded to the logging stride so that every slice samples the same rays as the
  // whole bundle would (ray i is logged iff i % record_every == 0)
  std::vector<long long> cut(n_devices + 1);
  for (int d = 0; d <= n_devices; ++d) {
    long long at = n * d / n_devices;
    if (every > 1 && d < n_devices) at = (at + every - 1) / every * every;
    cut[d] = at < n ? at : n;
  }
  cut[n_devices] = n;
  struct Part {
    std::vector<int64_t> distinct, cross, bins, stats;
    std::vector<double> sums;
    int32_t no_counts = 0;
    double elapsed = 0.0;
    int rc = 0;
    std::string error;
  };
  std::vector<Part> parts(n_devices);
  std::vector<std::thread> threads;
  for (int d = 0; d < n_devices; ++d) {
    Part& part = parts[d];
    part.distinct.assign(R > 0 ? R : 1, 0); part.cross.assign(R > 0 ? R : 1, 0); part.sums.assign(R > 0 ? 8 * R : 1, 0.0);
    part.bins.assign(B > 0 ? B : 1, 0); part.stats.assign(PVT_NSTATS, 0);
    threads.emplace_back([&, d] {
      Part& mine = parts[d];
      const long long lo = cut[d], m = cut[d + 1] - cut[d];
      if (m <= 0) return;
      pvt_params_t p = *params;
      p.n = m; p.first_index = params->first_index + lo; p.device = device_ids[d];
      pvt_out_t o = *out;
      o.rec_distinct = mine.distinct.data(); o.rec_crossings = mine.cross.data(); o.rec_sums = mine.sums.data();
      o.rec_bins = mine.bins.data(); o.stats = mine.stats.data();
      if (every > 0 && out->kind) {  // the slice's log rows, in place
        const long long ray0 = lo / every, row0 = ray0 * params->max_events;
        o.counts = out->counts + ray0; o.kind = out->kind + row0;
        o.hit = out->hit + row0; o.container = out->container + row0; o.adjacent = out->adjacent + row0;
        o.component = out->component + row0; o.source = out->source + row0;
        o.position = out->position + 3 * row0; o.direction = out->direction + 3 * row0; o.normal = out->normal + 3 * row0;
        o.wavelength = out->wavelength + row0; o.travelled = out->travelled + row0; o.duration = out->duration + row0;
      } else {
        o.counts = &mine.no_counts;
      }
      mine.rc = trace_host_bundle(scene, emit, have_rays ? positions + 3 * lo : nullptr, have_rays ? directions + 3 * lo : nullptr,
                                  have_rays ? wavelengths + lo : nullptr, &p, &o, &mine.elapsed);
      if (mine.rc) mine.error = pvt_last_error();  // (the message is thread-local)
    });
  }
  for (auto& t : threads) t.join();
  for (int d = 0; d < n_devices; ++d)
    if (parts[d].rc) return fail("device %d: %s", device_ids[d], parts[d].error.c_str());
  for (int r = 0; r < R; ++r) { out->rec_distinct[r] = 0; out->rec_crossings[r] = 0; }
  for (int k = 0; k < 8 * R; ++k) out->rec_sums[k] = 0.0;
  for (int b = 0; b < B; ++b) out->rec_bins[b] = 0;
  if (out->stats) for (int k = 0; k < PVT_NSTATS; ++k) out->stats[k] = 0;
  double slowest = 0.0;
  for (int d = 0; d < n_devices; ++d) {
    const Part& part = parts[d];
    for (int r = 0; r < R; ++r) { out->rec_distinct[r] += part.distinct[r]; out->rec_crossings[r] += part.cross[r]; }
    for (int k = 0; k < 8 * R; ++k) out->rec_sums[k] += part.sums[k];
    for (int b = 0; b < B; ++b) out->rec_bins[b] += part.bins[b];
    if (out->stats) for (int k = 0; k < PVT_NSTATS; ++k) out->stats[k] += part.stats[k];
    slowest = part.elapsed > slowest ? part.elapsed : slowest;
  }
  if (elapsed_s) *elapsed_s = slowest;
  return 0;
}

extern "C" int pvt_emit_bundle(const pvt_emit_t* emit, double* positions, double* directions, double* wavelengths, int64_t n,
                               int64_t first_index, uint64_t seed, int device) {
  if (!emit || emit->n_lights <= 0) return fail("emitter has no lights");
  if (n <= 0) return 0;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return fail("no CUDA device is usable (pvtrace_b200 has no CPU fallback)");
  PVT_CUDA(cudaSetDevice(device));
  // an emitter-only blob: a scene with no nodes
  pvt_scene_t empty;
  memset(&empty, 0, sizeof(empty));
  std::vector<double> blob = pack_scene(empty, emit);
  DeviceBuffer<double> d_blob, d_rays;
  int rc = d_blob.reserve(blob.size());
  if (!rc) rc = d_rays.reserve((size_t)7 * n);
  if (!rc && cudaMemcpy(d_blob.ptr, blob.data(), blob.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess)
    rc = fail("emitter upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (!rc) {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    emit_kernel<<<grid_for(n, prop.multiProcessorCount), 256>>>(d_blob.ptr, d_rays.ptr, d_rays.ptr + 3 * n, d_rays.ptr + 6 * n, n,
                                                              first_index, make_run_seed(seed));
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(positions, d_rays.ptr, 24 * (size_t)n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(directions, d_rays.ptr + 3 * n, 24 * (size_t)n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(wavelengths, d_rays.ptr + 6 * n, 8 * (size_t)n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail("emit failed: %s", cudaGetErrorString(e));
  }
  d_blob.release();
  d_rays.release();
  return rc;
}

extern "C" int pvt_intersect_bundle(const pvt_scene_t* scene, const double* positions, const double* directions, int64_t n,
                                    double* t0, int32_t* hit, int32_t* container, int32_t* adjacent, int device,
                                    double* elapsed_s) {
  if (n < 0) return fail("n must be >= 0");
  HostPath* path = nullptr;
  PVT_TRY(acquire_path(device, &path));
  std::lock_guard<std::mutex> lock(path->mutex);
  PVT_CUDA(cudaSetDevice(device));
  pvt_context* c = nullptr;
  PVT_TRY(acquire_context(*path, scene, nullptr, device, &c));
  PVT_CUDA(cudaSetDevice(c->device));
  if (n == 0) return 0;
  DeviceBuffer<double> in, d_t0;
  DeviceBuffer<int32_t> ids;
  int rc = in.reserve((size_t)6 * n);
  if (!rc) rc = d_t0.reserve((size_t)n);
  if (!rc) rc = ids.reserve((size_t)3 * n);
  Timer timer;
  if (!rc) rc = timer.start();
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(in.ptr, positions, 24 * (size_t)n, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(in.ptr + 3 * n, directions, 24 * (size_t)n, cudaMemcpyHostToDevice, 0);
    if (e != cudaSuccess) rc = fail("ray upload failed: %s", cudaGetErrorString(e));
  }
  if (!rc) rc = pvt_intersect_device(c, in.ptr, in.ptr + 3 * n, n, d_t0.ptr, ids.ptr, ids.ptr + n, ids.ptr + 2 * n, 0);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(t0, d_t0.ptr, 8 * (size_t)n, cudaMemcpyDeviceToHost, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hit, ids.ptr, 4 * (size_t)n, cudaMemcpyDeviceToHost, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(container, ids.ptr + n, 4 * (size_t)n, cudaMemcpyDeviceToHost, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(adjacent, ids.ptr + 2 * n, 4 * (size_t)n, cudaMemcpyDeviceToHost, 0);
    if (e != cudaSuccess) rc = fail("result download failed: %s", cudaGetErrorString(e));
  }
  if (!rc) rc = timer.stop(elapsed_s);
  in.release(); d_t0.release(); ids.release();
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Library queries

extern "C" int pvt_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail("tflops is NULL");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return fail("no CUDA device is usable (pvtrace_b200 has no CPU fallback)");
  PVT_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PVT_CUDA(cudaGetDeviceProperties(&prop, device));
  const int grid = prop.multiProcessorCount * 8, iterations = 1 << 14;
  DeviceBuffer<double> out;
  PVT_TRY(out.reserve((size_t)grid * 256));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // (the first warms up)
    Timer timer;
    PVT_TRY(timer.start());
    fp64_peak_kernel<<<grid, 256>>>(out.ptr, iterations);
    PVT_CUDA(cudaGetLastError());
    double seconds = 0.0;
    PVT_TRY(timer.stop(&seconds));
    const double rate = 2.0 * 8.0 * iterations * (double)grid * 256.0 / seconds * 1e-12;
    if (rep > 0 && rate > best) best = rate;
  }
  *tflops = best;
  return 0;
}

extern "C" int pvt_version(void) { return PVT_VERSION; }
extern "C" int pvt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
extern "C" const char* pvt_last_error(void) { return error_buffer(); }
extern "C" void pvt_struct_sizes(int32_t sizes[4]) {
  sizes[0] = (int32_t)sizeof(pvt_scene_t); sizes[1] = (int32_t)sizeof(pvt_emit_t);
  sizes[2] = (int32_t)sizeof(pvt_params_t); sizes[3] = (int32_t)sizeof(pvt_out_t);
}

// ---------------------------------------------------------------------------------------------------------
// Known-answer helpers: upload, one launch, download.

namespace {
struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() { for (void* p : ptrs) cudaFree(p); }
  template <class T>
  T* up(const T* host, size_t count) {
    T* d = nullptr;
    if (cudaMalloc((void**)&d, (count ? count : 1) * sizeof(T)) != cudaSuccess) return nullptr;
    ptrs.push_back(d);
    if (host && count && cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    return d;
  }
};
int begin_test(int device) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return fail("no CUDA device is usable (pvtrace_b200 has no CPU fallback)");
  PVT_CUDA(cudaSetDevice(device));
  return 0;
}
template <class T>
int finish_test(T* host, const T* dev, size_t count) {
  PVT_CUDA(cudaGetLastError());
  PVT_CUDA(cudaMemcpy(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost));
  return 0;
}
inline int blocks(int64_t n) { return (int)((n + 255) / 256); }
}  // namespace

#define PVT_NEED(p) if (!(p)) return fail("device allocation/upload failed")

extern "C" int pvt_test_fresnel_reflectivity(int64_t n, const double* angle, const double* n1, const double* n2, double* out, int device) {
  PVT_TRY(begin_test(device));
  if (n <= 0) return 0;
  Scratch s;
  double *a = s.up(angle, n), *b = s.up(n1, n), *c = s.up(n2, n), *o = s.up<double>(nullptr, n);
  PVT_NEED(a && b && c && o);
  test_fresnel_kernel<<<blocks(n), 256>>>(n, a, b, c, o);
  return finish_test(out, o, n);
}
extern "C" int pvt_test_specular_reflect(int64_t n, const double* d, const double* nrm, double* out, int device) {
  PVT_TRY(begin_test(device));
  if (n <= 0) return 0;
  Scratch s;
  double *a = s.up(d, 3 * n), *b = s.up(nrm, 3 * n), *o = s.up<double>(nullptr, 3 * n);
  PVT_NEED(a && b && o);
  test_reflect_kernel<<<blocks(n), 256>>>(n, a, b, o);
  return finish_test(out, o, 3 * n);
}
extern "C" int pvt_test_fresnel_refract(int64_t n, const double* d, const double* nrm, const double* n1, const double* n2, double* out, int device) {
  PVT_TRY(begin_test(device));
  if (n <= 0) return 0;
  Scratch s;
  double *a = s.up(d, 3 * n), *b = s.up(nrm, 3 * n), *c = s.up(n1, n), *e = s.up(n2, n), *o = s.up<double>(nullptr, 3 * n);
  PVT_NEED(a && b && c && e && o);
  test_refract_kernel<<<blocks(n), 256>>>(n, a, b, c, e, o);
  return finish_test(out, o, 3 * n);
}
extern "C" int pvt_test_intersect(int64_t n, const int32_t* geom_type, const double* params, const double* o, const double* d,
                                  int32_t* nhit, double* ts, int device) {
  PVT_TRY(begin_test(device));
  if (n <= 0) return 0;
  Scratch s;
  int32_t* g = s.up(geom_type, n);
  double *p = s.up(params, 4 * n), *oo = s.up(o, 3 * n), *dd = s.up(d, 3 * n), *t = s.up<double>(nullptr, 4 * n);
  int32_t* k = s.up<int32_t>(nullptr, n);
  PVT_NEED(g && p && oo && dd && t && k);
  test_intersect_kernel<<<blocks(n), 256>>>(n, g, p, oo, dd, k, t);
  PVT_TRY(finish_test(nhit, k, n));
  return finish_test(ts, t, 4 * n);
}
extern "C" int pvt_test_local_normal(int64_t n, const int32_t* geom_type, const double* params, const double* p, double* out, int device) {
  PVT_TRY(begin_test(device));
  if (n <= 0) return 0;
  Scratch s;
  int32_t* g = s.up(geom_type, n);
  double *q = s.up(params, 4 * n), *pp = s.up(p, 3 * n), *o = s.up<double>(nullptr, 3 * n);
  PVT_NEED(g && q && pp && o);
  test_normal_kernel<<<blocks(n), 256>>>(n, g, q, pp, o);
  return finish_test(out, o, 3 * n);
}
extern "C" int pvt_test_interp(int64_t n, const double* x, int32_t m, const double* xs, const double* ys, double* out, int device) {
  PVT_TRY(begin_test(device));
  if (n <= 0) return 0;
  if (m < 1) return fail("table needs at least one knot");
  Scratch s;
  double *a = s.up(x, n), *b = s.up(xs, m), *c = s.up(ys, m), *o = s.up<double>(nullptr, n);
  PVT_NEED(a && b && c && o);
  // same host decision as the scene packer: uniform grids take the guessed-bracket path
  test_interp_kernel<<<blocks(n), 256>>>(n, a, m, b, c, uniform_inv_dx(xs, m), o);
  return finish_test(out, o, n);
}
extern "C" int pvt_test_math(int64_t n, int32_t op, const double* a, const double* b, double* out, int device) {
  PVT_TRY(begin_test(device));
  if (n <= 0) return 0;
  if (op < 0 || op > 3) return fail("pvt_test_math: op must be 0 (log), 1 (divide), 2 (reciprocal) or 3 (sqrt)");
  Scratch s;
  double *da = s.up(a, n), *db = s.up(b ? b : a, n), *o = s.up<double>(nullptr, n);
  PVT_NEED(da && db && o);
  test_math_kernel<<<blocks(n), 256>>>(n, op, da, db, o);
  return finish_test(out, o, n);
}
extern "C" int pvt_test_rng_uniform(int64_t n_rays, int32_t n_draws, uint64_t seed, int64_t first_index, int32_t rng_mode, double* out, int device) {
  PVT_TRY(begin_test(device));
  if (n_rays <= 0 || n_draws <= 0) return 0;
  Scratch s;
  double* o = s.up<double>(nullptr, (size_t)n_rays * n_draws);
  PVT_NEED(o);
  if (rng_mode == PVT_RNG_XOSHIRO) test_rng_kernel<XoshiroStream><<<blocks(n_rays), 256>>>(n_rays, n_draws, seed, first_index, o);
  else test_rng_kernel<PhiloxStream><<<blocks(n_rays), 256>>>(n_rays, n_draws, seed, first_index, o);
  return finish_test(out, o, (size_t)n_rays * n_draws);
}
extern "C" int pvt_test_sample_phase(int64_t n, int32_t phase_type, double phase_param, uint64_t seed, int32_t rng_mode, double* out, int device) {
  PVT_TRY(begin_test(device));
  if (n <= 0) return 0;
  Scratch s;
  double* o = s.up<double>(nullptr, 3 * n);
  PVT_NEED(o);
  if (rng_mode == PVT_RNG_XOSHIRO) test_phase_kernel<XoshiroStream><<<blocks(n), 256>>>(n, phase_type, phase_param, seed, o);
  else test_phase_kernel<PhiloxStream><<<blocks(n), 256>>>(n, phase_type, phase_param, seed, o);
  return finish_test(out, o, 3 * n);
}
