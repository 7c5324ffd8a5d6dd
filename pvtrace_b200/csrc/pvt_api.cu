// pvt_api.cu -- the C ABI of include/pvtrace_b200.h on top of the kernels in pvt_kernels.cuh.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC  (csrc/build.py)
// There is no CPU path in this library: every compute entry fails with an error string if CUDA cannot run it.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
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
  // warp_wavefront_kernel (autonomous warps): shape chosen for the scene, 0 warps: does not fit
  int wave2_warps = 0, wave2_slots = 0;
  bool prefer_wave2 = false;
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
      size_t need = wavefront_smem_bytes(c->blob_words, pl);
      if (const char* env = getenv("PVT_EXTRA_SMEM")) need += (size_t)atoi(env);  // experiment: shrink L1
      if ((need + 1024) * b <= (size_t)prop.sharedMemPerMultiprocessor && need <= (size_t)prop.sharedMemPerBlockOptin) {
        c->wave_threads = t; c->wave_pool = pl; c->wave_ctas = b; c->wave_smem = need;
        break;
      }
    }
  }

  if (c->R() <= 64) {
    int want_w = 0, want_n = 0;
    if (const char* env = getenv("PVT_WAVE2_WARPS")) want_w = atoi(env);
    if (const char* env = getenv("PVT_WAVE2_SLOTS")) want_n = atoi(env);
    for (int k = 0; k < wave2_variant_count(); ++k) {
      const Wave2Variant v = wave2_variant(k);
      if ((want_w && v.warps != want_w) || (want_n && v.slots != want_n)) continue;
      const size_t need = wave2_smem(v, c->blob_words, true);
      if (need + 1024 <= (size_t)prop.sharedMemPerMultiprocessor && need <= (size_t)prop.sharedMemPerBlockOptin) {
        c->wave2_warps = v.warps; c->wave2_slots = v.slots;
        break;
      }
    }
  }
  if (const char* env = getenv("PVT_KERNEL")) c->prefer_wave2 = strcmp(env, "wave2") == 0;

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
  if (!rc && c->wave2_warps > 0)
    rc = wave2_setup(Wave2Variant{c->wave2_warps, c->wave2_slots}, wave2_smem(Wave2Variant{c->wave2_warps, c->wave2_slots}, c->blob_words, true));
  if (!rc && c->smem_bytes > 48 * 1024 &&
      (cudaFuncSetAttribute(intersect_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes) != cudaSuccess ||
       cudaFuncSetAttribute(intersect_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes) != cudaSuccess ||
       cudaFuncSetAttribute(intersect_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes) != cudaSuccess ||
       cudaFuncSetAttribute(intersect_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes) != cudaSuccess ||
       cudaFuncSetAttribute(intersect_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes) != cudaSuccess ||
       cudaFuncSetAttribute(intersect_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes) != cudaSuccess))
    rc = fail("cudaFuncSetAttribute(intersect_kernel) failed");
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
  c->blob.release(); c->tallies.release(); c->packed.release(); c->slabs.release(); c->arrived.release();
  c->counts.release(); c->hit.release(); c->container.release(); c->adjacent.release(); c->component.release();
  c->source.release(); c->kind.release(); c->position.release(); c->direction.release(); c->normal.release();
  c->wavelength.release(); c->travelled.release(); c->duration.release();
  delete c;
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

// does this bundle go through warp_wavefront_kernel?  (same conditions as the CTA wavefront; chosen by PVT_KERNEL=wave2
// or the PVT_FLAG_WARP_KERNEL bit)
static bool use_wave2(const pvt_context* c, const pvt_params_t* P) {
  if (c->wave2_warps <= 0 || P->rng_mode != PVT_RNG_PHILOX || (P->flags & PVT_FLAG_REGISTER_KERNEL) || P->n <= 0) return false;
  if (P->n >= (1ll << 32) - 1024) return false;
  if (P->flags & PVT_FLAG_CTA_KERNEL) return false;
  return c->prefer_wave2 || (P->flags & PVT_FLAG_WARP_KERNEL) || c->wave_threads <= 0;
}
static int wave2_grid(const pvt_context* c, const pvt_params_t* P) {
  const long long per_cta = (long long)c->wave2_warps * c->wave2_slots;
  const long long want_blocks = (P->n + per_cta - 1) / per_cta;
  return (int)(want_blocks < c->sm_count ? want_blocks : c->sm_count);
}

static int trace_device_impl(pvt_context_t* c, const double* d_pos, const double* d_dir, const double* d_wl,
                             const pvt_params_t* P, void* stream, const uint32_t* arrived);

extern "C" int pvt_trace_device(pvt_context_t* c, const double* d_pos, const double* d_dir, const double* d_wl,
                                const pvt_params_t* P, void* stream) {
  return trace_device_impl(c, d_pos, d_dir, d_wl, P, stream, nullptr);
}

static int trace_device_impl(pvt_context_t* c, const double* d_pos, const double* d_dir, const double* d_wl,
                             const pvt_params_t* P, void* stream, const uint32_t* arrived) {
  if (!c) return fail("ctx is NULL");
  PVT_TRY(check_params(P));
  const bool have_rays = d_pos && d_dir && d_wl;
  if (!have_rays && !c->has_emitter) return fail("no ray arrays given and the context has no emitter");
  PVT_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  PVT_TRY(prepare_log(c, P, st));
  if (P->n == 0) return 0;
  PVT_CUDA(cudaMemsetAsync(c->d_work(), 0, 8, st));

  TraceArgs a;
  a.hdr = c->hdr;
  a.blob = c->blob.ptr; a.blob_words = c->blob_words; a.scene_in_smem = c->scene_in_smem;
  a.pos = have_rays ? d_pos : nullptr; a.dir = have_rays ? d_dir : nullptr; a.wl = have_rays ? d_wl : nullptr;
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
  if (use_wave2(c, P)) {
    grid = wave2_grid(c, P);
    const Wave2Variant v{c->wave2_warps, c->wave2_slots};
    PVT_CUDA(cudaMemsetAsync(c->slabs.ptr, 0, (size_t)grid * 10 * c->R() * 8 + 8, st));
    PVT_TRY(wave2_launch(v, c->wave_boxes, P->record_every > 0, a, grid, wave2_smem(v, c->blob_words, P->record_every > 0), st));
  } else if (grid > 0) {
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

// resident CTAs of intersect_kernel (3 per SM by its launch bounds, fewer if shared memory says so)
template <class K>
static int intersect_grid(const pvt_context* c, long long n, K kernel) {
  int per_sm = 3;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, c->smem_bytes);
  if (per_sm < 1) per_sm = 1;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)c->sm_count * per_sm;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
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
  int ctas = 3;  // 80 registers with one tile in flight: measured best (4.26 TB/s of 68 B/ray traffic on config 2)
  if (const char* env = getenv("PVT_INTERSECT_CTAS")) ctas = atoi(env);
#define PVT_INTERSECT_LAUNCH(K, BX)                                                                                  \
  intersect_kernel<K, BX><<<intersect_grid(c, n, intersect_kernel<K, BX>), 256, c->smem_bytes, (cudaStream_t)stream>>>( \
      c->hdr, c->blob.ptr, c->blob_words, c->scene_in_smem, d_pos, d_dir, n, d_t0, d_hit, d_container, d_adjacent)
  if (c->wave_boxes) {
    if (ctas <= 2) PVT_INTERSECT_LAUNCH(2, true);
    else if (ctas == 3) PVT_INTERSECT_LAUNCH(3, true);
    else PVT_INTERSECT_LAUNCH(4, true);
  } else {
    if (ctas <= 2) PVT_INTERSECT_LAUNCH(2, false);
    else if (ctas == 3) PVT_INTERSECT_LAUNCH(3, false);
    else PVT_INTERSECT_LAUNCH(4, false);
  }
  PVT_CUDA(cudaGetLastError());
  return 0;
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
constexpr size_t kStreamChunkRays = 1 << 16;  // smallest chunk of the streaming upload (one launch, arrival marks)

// The streaming upload needs the copy stream to make progress WHILE the trace kernel runs.  Anything that
// serialises the two (CUDA_LAUNCH_BLOCKING, a profiler replaying kernels one at a time) would leave the kernel
// polling until it gives up, so the path is switched off when such a tool is detected, when the user says so
// (PVT_STREAM_UPLOAD=0), and for good after a bundle that did not complete (which is then re-traced the plain way).
bool g_stream_upload_ok = true;
bool stream_upload_allowed() {
  if (!g_stream_upload_ok) return false;
  if (const char* env = getenv("PVT_STREAM_UPLOAD")) return atoi(env) != 0;
  if (const char* env = getenv("CUDA_LAUNCH_BLOCKING")) { if (atoi(env) != 0) return false; }
  if (getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NVTX_INJECTION64_PATH"))
    return false;
  return true;
}

// One cached context per process: bundles of the same scene (engine.simulate_stream, repeated simulate calls)
// reuse the uploaded blob and every device buffer.
std::mutex g_cache_mutex;
pvt_context* g_cached = nullptr;
DeviceBuffer<double> g_rays;  // staging for host rays: [pos 3n | dir 3n | wl n]
int g_rays_device = -1;

int acquire_context(const pvt_scene_t* scene, const pvt_emit_t* emit, int device, pvt_context** out) {
  PVT_TRY(validate_scene(scene));
  std::vector<double> blob = pack_scene(*scene, emit);
  if (g_cached && g_cached->device == device && g_cached->host_blob.size() == blob.size() &&
      memcmp(g_cached->host_blob.data(), blob.data(), blob.size() * 8) == 0 && g_cached->has_emitter == (emit != nullptr)) {
    *out = g_cached;
    return 0;
  }
  if (g_cached) { pvt_context_destroy(g_cached); g_cached = nullptr; }
  PVT_TRY(pvt_context_create(scene, emit, device, &g_cached));
  *out = g_cached;
  return 0;
}
}  // namespace

extern "C" int pvt_trace_bundle(const pvt_scene_t* scene, const pvt_emit_t* emit, const double* positions,
                                const double* directions, const double* wavelengths, const pvt_params_t* params,
                                pvt_out_t* out, double* elapsed_s) {
  PVT_TRY(check_params(params));
  if (!out) return fail("out is NULL");
  const bool have_rays = positions && directions && wavelengths;
  if (!have_rays && !emit) return fail("either ray arrays or an emitter are required");
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  pvt_context* c = nullptr;
  PVT_TRY(acquire_context(scene, emit, params->device, &c));
  PVT_CUDA(cudaSetDevice(c->device));
  // two non-blocking streams: host rays are uploaded in chunks on one while the previous chunk is traced on the
  // other (overlap needs page-locked host memory; with pageable memory the copies simply serialise)
  static cudaStream_t s_copy = nullptr, s_run = nullptr;
  static cudaEvent_t s_uploaded[kMaxChunks] = {nullptr};
  static int s_device = -1;
  if (s_device != c->device) {
    if (s_copy) { cudaStreamDestroy(s_copy); cudaStreamDestroy(s_run); for (auto& e : s_uploaded) cudaEventDestroy(e); }
    PVT_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    PVT_CUDA(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
    for (auto& e : s_uploaded) PVT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s_device = c->device;
  }
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  PVT_CUDA(cudaEventCreate(&t0));
  PVT_CUDA(cudaEventCreate(&t1));
  PVT_CUDA(cudaEventRecord(t0, s_run));
  int rc = pvt_context_reset(c, s_run);
  const size_t n = (size_t)params->n;
  // Opt-in (PVT_ZERO_COPY=1): page-locked host arrays (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) are
  // read by the kernel IN PLACE, the warps that fill the shared-memory ray ring pulling them over PCIe a whole ring
  // ahead of their use.  Measured 42-45 GB/s against the copy engine's 55 GB/s, so the streaming upload below
  // is the default.
  bool streamed = false;
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
  } else if (!rc && have_rays && n > 0 && (wave_grid(c, params) > 0 || use_wave2(c, params)) && stream_upload_allowed()) {
    // Streaming upload: the trace kernel starts at once and polls an arrival mark; the copy engine delivers the
    // arrays front to back in chunks (three plain copies each) followed, in stream order, by the new mark = rays
    // complete so far.  CTAs claim photons in index order, so they consume the prefix as it lands.  Upload and trace
    // overlap completely: total time ~ max(PCIe, kernel) + the trace of the LAST chunk -- hence chunks that start
    // small (the kernel gets going at once), double up to a fifth of what is left (few, efficient copies) and
    // shrink again towards the end (little left to trace when the last byte lands).
    streamed = true;
    static uint32_t* h_marks = nullptr;  // page-locked: the mark copies must not be staged
    if (!h_marks) PVT_CUDA(cudaHostAlloc((void**)&h_marks, kMaxChunks * sizeof(uint32_t), cudaHostAllocDefault));
    if (g_rays_device != c->device) { g_rays.release(); g_rays_device = c->device; }
    rc = g_rays.reserve(7 * n);
    double *d_pos = g_rays.ptr, *d_dir = g_rays.ptr + 3 * n, *d_wl = g_rays.ptr + 6 * n;
    size_t min_chunk = kStreamChunkRays, shrink_div = 5;
    if (const char* env = getenv("PVT_UPLOAD_MIN_CHUNK")) { const long v = atol(env); if (v >= 1024) min_chunk = (size_t)v; }
    if (const char* env = getenv("PVT_UPLOAD_SHRINK")) { const int v = atoi(env); if (v >= 2 && v <= 64) shrink_div = (size_t)v; }
    int chunks = 0;
    cudaError_t e = cudaSuccess;
    static cudaEvent_t dbg[2] = {nullptr, nullptr};  // PVT_DEBUG_TIMING=1: how long the copy stream took
    const bool dbg_on = getenv("PVT_DEBUG_TIMING") != nullptr;
    if (dbg_on && !dbg[0]) { cudaEventCreate(&dbg[0]); cudaEventCreate(&dbg[1]); }
    if (dbg_on) cudaEventRecord(dbg[0], s_copy);
    if (!rc) e = cudaMemsetAsync(c->arrived.ptr, 0, 4, s_copy);
    if (!rc && e == cudaSuccess) e = cudaEventRecord(s_uploaded[0], s_copy);
    if (!rc && e == cudaSuccess) e = cudaStreamWaitEvent(s_run, s_uploaded[0], 0);
    if (!rc && e == cudaSuccess) rc = trace_device_impl(c, d_pos, d_dir, d_wl, params, s_run, c->arrived.ptr);
    size_t lo = 0, grow = min_chunk;
    while (lo < n && !rc && e == cudaSuccess) {
      const size_t left = n - lo;
      size_t m = left / shrink_div > min_chunk ? left / shrink_div : min_chunk;
      if (m > grow) m = grow;
      grow *= 2;
      if (m > left || chunks == kMaxChunks - 1 || left - m < min_chunk / 2) m = left;
      e = cudaMemcpyAsync(d_pos + 3 * lo, positions + 3 * lo, 24 * m, cudaMemcpyHostToDevice, s_copy);
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_dir + 3 * lo, directions + 3 * lo, 24 * m, cudaMemcpyHostToDevice, s_copy);
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_wl + lo, wavelengths + lo, 8 * m, cudaMemcpyHostToDevice, s_copy);
      lo += m;
      h_marks[chunks] = (uint32_t)lo;
      if (e == cudaSuccess) e = cudaMemcpyAsync(c->arrived.ptr, h_marks + chunks, 4, cudaMemcpyHostToDevice, s_copy);
      ++chunks;
    }
    if (dbg_on) {
      cudaEventRecord(dbg[1], s_copy);
      cudaEventSynchronize(dbg[1]);
      cudaStreamSynchronize(s_run);
      float up = 0.f, t_start = 0.f;
      cudaEventElapsedTime(&up, dbg[0], dbg[1]);
      cudaEventElapsedTime(&t_start, t0, dbg[0]);
      cudaEventRecord(t1, s_run); cudaEventSynchronize(t1);
      float total = 0.f;
      cudaEventElapsedTime(&total, t0, t1);
      fprintf(stderr, "[pvt] upload started %.3f ms after t0, took %.3f ms (%d chunks); trace done at %.3f ms\n", t_start, up, chunks, total);
    }
    if (!rc && e != cudaSuccess) {
      // never leave the kernel polling: publish "everything arrived" so it drains, then report
      h_marks[kMaxChunks - 1] = 0xffffffffu;
      cudaMemcpyAsync(c->arrived.ptr, h_marks + kMaxChunks - 1, 4, cudaMemcpyHostToDevice, s_copy);
      rc = fail("ray upload failed: %s", cudaGetErrorString(e));
    }
  } else if (!rc && have_rays && n > 0) {
    if (g_rays_device != c->device) { g_rays.release(); g_rays_device = c->device; }
    rc = g_rays.reserve(7 * n);
    double *d_pos = g_rays.ptr, *d_dir = g_rays.ptr + 3 * n, *d_wl = g_rays.ptr + 6 * n;
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
      if (e == cudaSuccess) e = cudaEventRecord(s_uploaded[k], s_copy);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s_run, s_uploaded[k], 0);
      if (e != cudaSuccess) { rc = fail("ray upload failed: %s", cudaGetErrorString(e)); break; }
      pvt_params_t part = *params;
      part.n = (int64_t)m;
      part.first_index = params->first_index + (int64_t)lo;
      rc = pvt_trace_device(c, d_pos + 3 * lo, d_dir + 3 * lo, d_wl + lo, &part, s_run);
    }
  } else if (!rc) {
    rc = pvt_trace_device(c, nullptr, nullptr, nullptr, params, s_run);
  }
  if (!rc) rc = pvt_context_read(c, out, s_run);
  if (!rc && n > 0) {  // every ray must have been traced
    u64 traced = 0;
    if (cudaMemcpy(&traced, c->d_stats() + PVT_STAT_RAYS, 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
      rc = fail("reading the ray counter failed: %s", cudaGetErrorString(cudaGetLastError()));
    } else if (traced != (u64)n && streamed) {
      // the kernel gave up polling for rays: something serialised upload and trace.  Trace again the plain way.
      g_stream_upload_ok = false;
      cudaStreamSynchronize(s_copy);
      rc = pvt_context_reset(c, s_run);
      if (!rc) rc = pvt_trace_device(c, g_rays.ptr, g_rays.ptr + 3 * n, g_rays.ptr + 6 * n, params, s_run);
      if (!rc) rc = pvt_context_read(c, out, s_run);
    } else if (traced != (u64)n) {
      rc = fail("traced %llu of %zu rays", traced, n);
    }
  }
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
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  pvt_context* c = nullptr;
  PVT_TRY(acquire_context(scene, nullptr, device, &c));
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
