// pvt_kernels.cuh -- the CUDA kernels of libpvtrace_b200.so (sm_100a).
//
//   trace_kernel      persistent-threads photon tracer.  One CTA per SM slot for the whole launch; the scene blob
//                     is staged into shared memory once per CTA by a single TMA bulk copy (cp.async.bulk +
//                     mbarrier); every lane owns one live photon in registers and advances it one step per loop
//                     iteration; lanes whose photon retired are found by warp ballot and refilled from a
//                     warp-private reservoir of photon indices, itself refilled from one global counter.  All
//                     randomness is counter based (photon index, draw number), so the schedule is invisible in
//                     the results.
//   intersect_kernel  the ray/primitive stage on its own over a photon array (next_hit + find_container).
//   emit_kernel       initial rays of the built-in light delegates.
//   *_test kernels    one thin launch per device helper for the known-answer tests.
#pragma once
#include <cuda_runtime.h>

#include "pvt_photon.cuh"

namespace pvt {

constexpr int kTraceThreads = 256;
constexpr int kReservoirChunk = 64;  // photon indices a warp takes from the global counter at a time (>= 32)
constexpr unsigned kFullMask = 0xffffffffu;

struct TraceArgs {
  const double* blob;  // scene blob in global memory
  int blob_words;
  int scene_in_smem;   // 0: the blob did not fit into shared memory, read it through L1/L2 instead
  const double* pos;   // [n,3] or null (=> emit on device)
  const double* dir;
  const double* wl;
  long long n, first_index, record_every;
  u64 seed;
  StepParams sp;
  u64* work_counter;
  u64* g_distinct;  // [R]
  u64* g_cross;     // [R]
  double* g_sums;   // [R,8]
  u64* g_bins;      // [total_bins]
  u64* g_stats;     // [PVT_NSTATS]
  LogColumns log;
};

// ---- shared-memory staging of the scene blob through the TMA engine ---------------------------------------

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Copies `bytes` (multiple of 16, both sides 16-byte aligned) global -> shared with one bulk async copy issued
// by thread 0 and waits for it on an mbarrier.  Ends with every thread of the CTA able to read the data.
__device__ __forceinline__ void stage_blob(double* dst, const double* src, uint32_t bytes, uint64_t* bar) {
  const uint32_t bar_a = smem_addr(bar), dst_a = smem_addr(dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
                 "l"(src), "r"(bytes), "r"(bar_a)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar_a)
        : "memory");
  }
}

// shared memory layout of trace_kernel: [mbarrier 16 B][blob][slab: distinct R | cross R | sums 8R]
__host__ __device__ inline size_t trace_smem_bytes(int blob_words_in_smem, int n_recorders) {
  return 16 + (size_t)blob_words_in_smem * 8 + (size_t)n_recorders * 10 * 8;
}

template <class Rng, int SW>
__global__ void __launch_bounds__(kTraceThreads) trace_kernel(const TraceArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sblob = reinterpret_cast<double*>(smem_raw + 16);
  const int words_in_smem = a.scene_in_smem ? a.blob_words : 0;
  u64* slab = reinterpret_cast<u64*>(smem_raw + 16 + (size_t)words_in_smem * 8);

  if (a.scene_in_smem) stage_blob(sblob, a.blob, (uint32_t)a.blob_words * 8u, bar);
  const SceneView sv{a.scene_in_smem ? sblob : a.blob};
  const int R = sv.hdr().n_recorders;
  for (int k = threadIdx.x; k < R * 10; k += blockDim.x) slab[k] = 0ull;
  __syncthreads();
  const TallySink T{slab, slab + R, reinterpret_cast<double*>(slab + 2 * R), a.g_bins};

  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  Photon<Rng, SW> ph;
  ph.nsteps = 0; ph.nevents = 0;
  ph.log_base = -1; ph.nlog = 0;
  bool alive = false, exhausted = false;
  long long res_next = 0, res_end = 0;  // warp-uniform reservoir [res_next, res_end) of photon indices
  uint32_t rays = 0;

  for (;;) {
    const unsigned need = __ballot_sync(kFullMask, !alive && !exhausted);
    if (need) {
      const int cnt = __popc(need), rank = __popc(need & lt_mask);
      const bool mine = (need >> lane) & 1u;
      long long idx = -1;
      const long long avail = res_end - res_next;
      if (avail >= cnt) {
        if (mine) idx = res_next + rank;
        res_next += cnt;
      } else {
        if (mine && rank < avail) idx = res_next + rank;
        const int rem = cnt - (int)avail;
        long long base = 0;
        if (lane == 0) base = (long long)atomicAdd(a.work_counter, (u64)kReservoirChunk);
        base = __shfl_sync(kFullMask, base, 0);
        res_next = base < a.n ? base : a.n;
        res_end = base + kReservoirChunk < a.n ? base + kReservoirChunk : a.n;
        const long long avail2 = res_end - res_next;
        if (mine && rank >= avail && rank - avail < avail2) idx = res_next + (rank - avail);
        res_next += rem < avail2 ? rem : avail2;
      }
      if (mine) {
        if (idx >= 0) {
          const u64 id = a.seed + (u64)a.first_index + (u64)idx;
          if (a.pos) {
            ph.p = V3{a.pos[3 * idx], a.pos[3 * idx + 1], a.pos[3 * idx + 2]};
            ph.d = V3{a.dir[3 * idx], a.dir[3 * idx + 1], a.dir[3 * idx + 2]};
            ph.wl = a.wl[idx];
          } else {
            emit_ray(sv, id, a.first_index + idx, ph.p, ph.d, ph.wl);
          }
          ph.rng.init(id);
          ph.log_base = -1;
          if (a.record_every > 0 && idx % a.record_every == 0) ph.log_base = (idx / a.record_every) * (long long)a.sp.max_events;
          begin_photon(ph, a.log, a.sp);
          alive = true;
          ++rays;
        } else {
          exhausted = true;  // the global counter is past n: nothing will ever arrive
        }
      }
    }
    if (!__any_sync(kFullMask, alive)) break;
    if (alive) {
      alive = step_photon(sv, T, a.log, a.sp, ph);
      if (!alive && ph.log_base >= 0) {  // a sampled ray publishes its event count when it retires
        a.log.counts[ph.log_base / a.sp.max_events] = ph.nlog;
        ph.log_base = -1;
      }
    }
  }

  // run statistics: one atomic per warp per counter
  u64 s_steps = ph.nsteps, s_events = ph.nevents, s_rays = rays;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s_steps += __shfl_down_sync(kFullMask, s_steps, off);
    s_events += __shfl_down_sync(kFullMask, s_events, off);
    s_rays += __shfl_down_sync(kFullMask, s_rays, off);
  }
  if (lane == 0) {
    atomicAdd(a.g_stats + PVT_STAT_STEPS, s_steps);
    atomicAdd(a.g_stats + PVT_STAT_EVENTS, s_events);
    atomicAdd(a.g_stats + PVT_STAT_RAYS, s_rays);
  }
  // flush the CTA's tally slab
  __syncthreads();
  for (int k = threadIdx.x; k < R * 10; k += blockDim.x) {
    if (k < R) { if (slab[k]) atomicAdd(a.g_distinct + k, slab[k]); }
    else if (k < 2 * R) { if (slab[k]) atomicAdd(a.g_cross + (k - R), slab[k]); }
    else {
      const double v = reinterpret_cast<double*>(slab)[k];
      if (v != 0.0) atomicAdd(a.g_sums + (k - 2 * R), v);
    }
  }
}

// ---- the intersect stage on its own ----------------------------------------------------------------------
// 60 B of algorithmic traffic per ray: reads position + direction (48 B), writes t0 (8 B) and three int32 ids
// (12 B; SURVEY 8d counts them packed as 4 B).
__global__ void __launch_bounds__(256) intersect_kernel(const double* blob, int blob_words, int scene_in_smem,
                                                        const double* pos, const double* dir, long long n, double* t0,
                                                        int32_t* hit, int32_t* container, int32_t* adjacent) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sblob = reinterpret_cast<double*>(smem_raw + 16);
  if (scene_in_smem) stage_blob(sblob, blob, (uint32_t)blob_words * 8u, bar);
  const SceneView sv{scene_in_smem ? sblob : blob};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const V3 p = V3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    const V3 d = V3{dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
    const Nearest nh = nearest_surface(sv, p, d);
    t0[i] = nh.total ? nh.t0 : PVT_INF;
    hit[i] = nh.total ? nh.hit : -1;
    container[i] = nh.container;
    adjacent[i] = nh.adjacent;
  }
}

__global__ void __launch_bounds__(256) emit_kernel(const double* blob, double* pos, double* dir, double* wl, long long n,
                                                   long long first_index, u64 seed) {
  const SceneView sv{blob};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    V3 p, d;
    double w;
    emit_ray(sv, seed + (u64)first_index + (u64)i, first_index + i, p, d, w);
    pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
    dir[3 * i] = d.x; dir[3 * i + 1] = d.y; dir[3 * i + 2] = d.z;
    wl[i] = w;
  }
}

// tallies <-> packed doubles (the buffer a multi-GPU caller all-reduces)
__global__ void pack_tallies_kernel(const u64* ints_a, int n_a, const double* sums, int n_s, const u64* bins, int n_b,
                                    double* packed) {
  const int total = n_a + n_s + n_b;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    double v;
    if (k < n_a) v = (double)ints_a[k];
    else if (k < n_a + n_s) v = sums[k - n_a];
    else v = (double)bins[k - n_a - n_s];
    packed[k] = v;
  }
}
__global__ void unpack_tallies_kernel(u64* ints_a, int n_a, double* sums, int n_s, u64* bins, int n_b,
                                      const double* packed) {
  const int total = n_a + n_s + n_b;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    const double v = packed[k];
    if (k < n_a) ints_a[k] = (u64)llrint(v);
    else if (k < n_a + n_s) sums[k - n_a] = v;
    else bins[k - n_a - n_s] = (u64)llrint(v);
  }
}

// ---- known-answer test kernels ----------------------------------------------------------------------------

__global__ void test_fresnel_kernel(long long n, const double* angle, const double* n1, const double* n2, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fresnel_R(angle[i], n1[i], n2[i]);
}
__global__ void test_reflect_kernel(long long n, const double* d, const double* nrm, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 r = mirror(V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]}, V3{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]});
  out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}
__global__ void test_refract_kernel(long long n, const double* d, const double* nrm, const double* n1, const double* n2,
                                    double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 dd = V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]};
  V3 nf = V3{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
  if (dot(nf, dd) < 0.0) nf = neg(nf);
  const V3 r = snell(dd, nf, n1[i], n2[i]);
  out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}
__global__ void test_intersect_kernel(long long n, const int32_t* gtype, const double* params, const double* o,
                                      const double* d, int32_t* nhit, double* ts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  const int k = roots(gtype[i], params + 4 * i, V3{o[3 * i], o[3 * i + 1], o[3 * i + 2]},
                      V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]}, t);
  nhit[i] = k;
  for (int j = 0; j < 4; ++j) ts[4 * i + j] = j < k ? t[j] : 0.0;
}
__global__ void test_normal_kernel(long long n, const int32_t* gtype, const double* params, const double* p, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 r = outward_normal(gtype[i], params + 4 * i, V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]});
  out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}
__global__ void test_interp_kernel(long long n, const double* x, int m, const double* xs, const double* ys, double inv_dx,
                                   double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = interp_hinted(x[i], xs, ys, m, inv_dx);
}
template <class Rng>
__global__ void test_rng_kernel(long long n_rays, int n_draws, u64 seed, long long first_index, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  Rng rng;
  rng.init(seed + (u64)first_index + (u64)i);
  for (int k = 0; k < n_draws; ++k) out[i * n_draws + k] = rng.next();
}
template <class Rng>
__global__ void test_phase_kernel(long long n, int ptype, double prm, u64 seed, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rng rng;
  rng.init(seed + (u64)i);
  const V3 r = phase_direction(ptype, prm, rng);
  out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}

}  // namespace pvt
