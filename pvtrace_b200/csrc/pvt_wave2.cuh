// pvt_wave2.cuh -- warp_wavefront_kernel: the tracer as AUTONOMOUS WARPS (sm_100a).
//
// The two-stage CTA wavefront (pvt_kernels.cuh) regroups photons by the kind of their next interaction once per
// iteration of the whole CTA: two CTA-wide barriers, two rounds of work stealing and a handful of shared atomics per
// step, and every photon's state crosses shared memory twice per step.  Here every WARP is a wavefront machine of its
// own.  It owns N photon slots (structure of arrays in shared memory) and four queues of slot numbers -- VOLUME,
// SURFACE, EXIT, FREE -- whose lengths live in one register.  One iteration of a warp:
//
//   1. take up to 32 entries of the fullest queue (FREE entries only as far as fresh rays are at hand);
//   2. INTERACT: all lanes run the same kind of event on their photon -- absorption / re-emission, reflection /
//      refraction, leaving the scene, or (FREE) taking the next ray of the bundle;
//   3. CLASSIFY the survivors at once, in the same registers: next_hit + find_container + Beer-Lambert free path;
//   4. append every slot to the queue of its next event with warp ballots.
//
// No barrier, no atomic and no counter is shared between warps (fresh rays are claimed from the bundle's global
// counter, 64 at a time): warps drift apart, so the schedulers always find warps in different phases (Philox integer
// chains next to fp64 slabs next to MUFU), and a photon's state crosses shared memory once per step.  Random numbers
// are addressed by (photon index, step, purpose): the results are bit-identical to those of the other kernels whatever
// the order in which photons advance.
#pragma once
#include "pvt_kernels.cuh"

namespace pvt {

enum { kQVolume = 0, kQSurface = 1, kQExit = 2, kQFree = 3 };
constexpr uint32_t kClaimWarp = 64;  // rays a warp claims from the bundle's counter at a time

// per-warp pool: 12 f64 columns (px py pz dx dy dz wl travelled duration | plan t, u, alpha), seen mask, three 32-bit
// columns (count, photon index, packed ids) + three more when events are logged (source, nlog, log_ray), four queues of
// N one-byte slot numbers
__host__ __device__ constexpr size_t warp_pool_bytes(int N, bool log) {
  return (size_t)N * (12 * 8 + 8 + 3 * 4 + (log ? 3 * 4 : 0) + 4);
}
__host__ __device__ inline size_t wave2_smem_bytes(int blob_words, int W, int N, bool log) {
  return 16 + (size_t)blob_words * 8 + (size_t)W * warp_pool_bytes(N, log);
}

template <int N, bool kLog>
struct WarpPool {
  double* f;
  u64* seen;
  int32_t* count;
  uint32_t *idx, *ids;
  int32_t *source, *nlog, *log_ray;
  uint8_t* q;
  __device__ __forceinline__ explicit WarpPool(unsigned char* base) {
    f = reinterpret_cast<double*>(base);
    seen = reinterpret_cast<u64*>(f + 12 * N);
    count = reinterpret_cast<int32_t*>(seen + N);
    idx = reinterpret_cast<uint32_t*>(count + N);
    ids = idx + N;
    int32_t* w = reinterpret_cast<int32_t*>(ids + N);
    if (kLog) { source = w; nlog = w + N; log_ray = w + 2 * N; w += 3 * N; }
    else { source = nlog = log_ray = nullptr; }
    q = reinterpret_cast<uint8_t*>(w);
  }
  __device__ __forceinline__ double& col(int c, int s) const { return f[c * N + s]; }
};

template <int N, bool kLog>
__device__ __forceinline__ void w2_load(const WarpPool<N, kLog>& pool, int s, PoolPhoton& ph, StepPlan& plan, uint32_t& ids,
                                        int max_events) {
  ph.p = V3{pool.col(0, s), pool.col(1, s), pool.col(2, s)};
  ph.d = V3{pool.col(3, s), pool.col(4, s), pool.col(5, s)};
  ph.wl = pool.col(6, s); ph.travelled = pool.col(7, s); ph.duration = pool.col(8, s);
  plan.t = pool.col(9, s); plan.u = pool.col(10, s); plan.alpha = pool.col(11, s);
  const u64 seen = pool.seen[s];
  ph.seen[0] = (uint32_t)seen; ph.seen[1] = (uint32_t)(seen >> 32);
  ph.count = pool.count[s];
  ids = pool.ids[s];
  plan.hit = (int)(ids & 0xff); plan.container = (int)((ids >> 8) & 0xff); plan.adjacent = (int)((ids >> 16) & 0xff);
  if (plan.adjacent == 0xff) plan.adjacent = -1;
  ph.source = -1; ph.nlog = 0; ph.log_ray = -1; ph.log_base = -1;
  if (kLog) {
    ph.source = pool.source[s]; ph.nlog = pool.nlog[s]; ph.log_ray = pool.log_ray[s];
    ph.log_base = ph.log_ray < 0 ? -1 : (long long)ph.log_ray * max_events;
  }
}

template <int N, bool kLog>
__device__ __forceinline__ void w2_store(const WarpPool<N, kLog>& pool, int s, const PoolPhoton& ph, const StepPlan& plan,
                                         StepClass cls) {
  pool.col(0, s) = ph.p.x; pool.col(1, s) = ph.p.y; pool.col(2, s) = ph.p.z;
  pool.col(3, s) = ph.d.x; pool.col(4, s) = ph.d.y; pool.col(5, s) = ph.d.z;
  pool.col(6, s) = ph.wl; pool.col(7, s) = ph.travelled; pool.col(8, s) = ph.duration;
  pool.col(9, s) = plan.t; pool.col(10, s) = plan.u; pool.col(11, s) = plan.alpha;
  pool.seen[s] = (u64)ph.seen[0] | ((u64)ph.seen[1] << 32);
  pool.count[s] = ph.count;
  pool.ids[s] = (uint32_t)(plan.hit & 0xff) | ((uint32_t)(plan.container & 0xff) << 8) |
                ((uint32_t)(plan.adjacent & 0xff) << 16) | (cls == kKill ? 1u << 24 : 0u);
  if (kLog) { pool.source[s] = ph.source; pool.nlog[s] = ph.nlog; }
}

// W warps per CTA (one CTA per SM), N slots per warp (multiple of 32, <= 224: queue lengths are bytes of one register)
template <int W, int N, bool kLog, bool kBoxes>
__global__ void __launch_bounds__(W * 32, 1) warp_wavefront_kernel(const __grid_constant__ TraceArgs a) {
  static_assert(N % 32 == 0 && N <= 224, "slots per warp");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sblob = reinterpret_cast<double*>(smem_raw + 16);
  stage_blob(sblob, a.blob, (uint32_t)a.blob_words * 8u, bar);
  const SceneView sv{sblob, &a.hdr};
  const int R = sv.hdr().n_recorders;
  const TallySink sink = cta_sink(a, R);
  const StepParams sp = a.sp;
  const LogColumns& L = a.log;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned below = (1u << lane) - 1u;
  const WarpPool<N, kLog> pool(smem_raw + 16 + (size_t)a.blob_words * 8 + (size_t)warp * warp_pool_bytes(N, kLog));
  for (int s = lane; s < N; s += 32) pool.q[kQFree * N + s] = (uint8_t)s;
  __syncwarp();

  // queue state, the same in every lane: lengths and heads of the four circular queues, one byte each
  uint32_t cnt = (uint32_t)N << (8 * kQFree), head = 0;
  // fresh rays: [res_next, res_end) of the bundle are this warp's to take; `mark` = rays that have arrived so far
  uint32_t res_next = 0, res_end = 0, mark = a.arrived ? 0u : 0xffffffffu, idle = 0;
  bool exhausted = false;
  LaneStats st;
#ifdef PVT_W2_STATS  // chunks and lanes per queue, cycles per phase (lane 0 of every warp) -> stats[8..]
  u64 w2_chunks[4] = {0, 0, 0, 0}, w2_lanes[4] = {0, 0, 0, 0}, w2_cyc[4] = {0, 0, 0, 0};
  long long w2_t = clock64();
#define PVT_W2_TICK(k) do { const long long now_ = clock64(); w2_cyc[k] += (u64)(now_ - w2_t); w2_t = now_; } while (0)
#else
#define PVT_W2_TICK(k) do { } while (0)
#endif

  for (;;) {
    const int nV = (int)(cnt & 0xffu), nS = (int)((cnt >> 8) & 0xffu), nE = (int)((cnt >> 16) & 0xffu), nF = (int)(cnt >> 24);
    if (nF > 0 && res_next == res_end && !exhausted) {
      u64 b = 0;
      if (lane == 0) b = atomicAdd(a.work_counter, (u64)kClaimWarp);
      b = __shfl_sync(kFullMask, b, 0);
      if (b >= (u64)a.n) {
        exhausted = true;
      } else {
        res_next = (uint32_t)b;
        const u64 left = (u64)a.n - b;
        res_end = res_next + (left < kClaimWarp ? (uint32_t)left : kClaimWarp);
      }
    }
    int rays = (int)(res_end - res_next);
    if (a.arrived && rays > 0) {  // streaming upload: only the prefix [0, mark) of the bundle is in device memory yet
      const uint32_t want = res_next + (uint32_t)(rays < 32 ? rays : 32);
      if (mark < want) mark = *reinterpret_cast<const volatile uint32_t*>(a.arrived);
      if (mark < res_end) rays = mark > res_next ? (int)(mark - res_next) : 0;
    }
    const int nFr = nF < rays ? nF : rays;
    // the fullest queue (a full chunk is a full chunk: 32 caps the comparison), ties in the order V, S, FREE, EXIT
    int q = kQVolume, best = nV < 32 ? nV : 32;
    { const int c = nS < 32 ? nS : 32; if (c > best) { best = c; q = kQSurface; } }
    { const int c = nFr < 32 ? nFr : 32; if (c > best) { best = c; q = kQFree; } }
    { const int c = nE < 32 ? nE : 32; if (c > best) { best = c; q = kQExit; } }
    if (best == 0) {
      if (res_next == res_end && exhausted) break;  // every slot is free and the bundle has been handed out
      if (++idle > kMaxIdleIterations) break;       // rays that never arrive must not hang the device
      __nanosleep(200);
      continue;
    }
    idle = 0;
    const int m = best;
#ifdef PVT_W2_STATS
    w2_chunks[q] += 1; w2_lanes[q] += (u64)m;
#endif
    PVT_W2_TICK(0);  // scheduling
    int slot = -1;
    {
      const int h = (int)((head >> (8 * q)) & 0xffu);
      if (lane < m) {
        int at = h + lane;
        at = at >= N ? at - N : at;
        slot = pool.q[q * N + at];
      }
      int h2 = h + m;
      h2 = h2 >= N ? h2 - N : h2;
      head = (head & ~(0xffu << (8 * q))) | ((uint32_t)h2 << (8 * q));
      cnt -= (uint32_t)m << (8 * q);
    }

    PoolPhoton ph;
    StepPlan plan;
    bool alive = false;
    u64 photon = 0;  // index of the lane's photon within the run
    if (q == kQFree) {
      if (slot >= 0) {
        const long long i = (long long)res_next + lane;
        photon = (u64)a.first_index + (u64)i;
        if (a.pos) {
          // L2-only loads: with a streaming upload a line cached in L1 could hold a neighbour that had not arrived
          ph.p = V3{__ldcg(a.pos + 3 * i), __ldcg(a.pos + 3 * i + 1), __ldcg(a.pos + 3 * i + 2)};
          ph.d = V3{__ldcg(a.dir + 3 * i), __ldcg(a.dir + 3 * i + 1), __ldcg(a.dir + 3 * i + 2)};
          ph.wl = __ldcg(a.wl + i);
        } else {
          const EmittedRay e = emit_ray_value(sv, a.keys, a.first_index + i);
          ph.p = e.pos; ph.d = e.dir; ph.wl = e.wl;
        }
        ph.log_ray = (kLog && a.record_every > 0) ? sampled_ordinal(i, a.record_every) : -1;
        ph.log_base = ph.log_ray < 0 ? -1 : (long long)ph.log_ray * sp.max_events;
        begin_photon<kLog>(ph, L, sp, st);
        pool.idx[slot] = (uint32_t)i;
        if (kLog) pool.log_ray[slot] = ph.log_ray;
        alive = true;
      }
      res_next += (uint32_t)m;
    } else if (slot >= 0) {
      uint32_t ids;
      w2_load(pool, slot, ph, plan, ids, sp.max_events);
      photon = (u64)a.first_index + (u64)pool.idx[slot];
      PhiloxStream rng;
      rng.init(a.keys, photon);
      rng.begin_step((uint32_t)ph.count);
      TallyReq tr;
      if (q == kQVolume) alive = volume_step<kLog>(sv, L, sp, ph, rng, st, plan, tr);
      else if (q == kQSurface) alive = surface_step<kLog>(sv, L, sp, ph, rng, st, plan, tr);
      else if (ids >> 24) kill_step<kLog>(sv, L, sp, ph, st, plan, tr);
      else exit_step<kLog>(sv, L, sp, ph, st, plan, tr);
      PVT_W2_TICK(1);  // interact
      if (tr.sel >= 0) tally(sv, sink, ph, tr);
    }
    __syncwarp();
    PVT_W2_TICK(2);  // tally (FREE chunks: fetching the rays)

    StepClass cls = kDead;
    if (alive) {
      PhiloxStream rng;
      rng.init(a.keys, photon);
      cls = classify_step<kLog, kBoxes>(sv, L, sp, ph, rng, st, plan);
      if (cls != kDead) w2_store(pool, slot, ph, plan, cls);
    }
    if (kLog && slot >= 0 && cls == kDead && ph.log_ray >= 0) L.counts[ph.log_ray] = ph.nlog;
    __syncwarp();
    PVT_W2_TICK(3);  // classify

    // every slot of the chunk goes to the queue of its next event
    {
      const int qq = cls == kVolume ? kQVolume : (cls == kSurface ? kQSurface : (cls == kDead ? kQFree : kQExit));
      const bool has = slot >= 0;
      const unsigned mv = __ballot_sync(kFullMask, has && qq == kQVolume), ms = __ballot_sync(kFullMask, has && qq == kQSurface),
                     me = __ballot_sync(kFullMask, has && qq == kQExit), mf = __ballot_sync(kFullMask, has && qq == kQFree);
      if (has) {
        const unsigned mine = qq == kQVolume ? mv : (qq == kQSurface ? ms : (qq == kQExit ? me : mf));
        int at = (int)((head >> (8 * qq)) & 0xffu) + (int)((cnt >> (8 * qq)) & 0xffu) + __popc(mine & below);
        at = at >= N ? at - N : at;
        pool.q[qq * N + at] = (uint8_t)slot;
      }
      cnt += (uint32_t)__popc(mv) | ((uint32_t)__popc(ms) << 8) | ((uint32_t)__popc(me) << 16) | ((uint32_t)__popc(mf) << 24);
      __syncwarp();
    }
  }
#ifdef PVT_W2_STATS
  if (lane == 0) {
    for (int k = 0; k < 4; ++k) {
      atomicAdd(a.g_stats + 8 + k, w2_chunks[k]); atomicAdd(a.g_stats + 12 + k, w2_lanes[k]); atomicAdd(a.g_stats + 16 + k, w2_cyc[k]);
    }
  }
#endif
  retire_cta(a, R, st);
}

}  // namespace pvt
