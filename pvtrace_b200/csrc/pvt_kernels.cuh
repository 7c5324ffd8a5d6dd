// pvt_kernels.cuh -- the CUDA kernels of libpvtrace_b200.so (sm_100a).
//
//   wavefront_kernel  THE tracer.  Persistent CTAs (one per SM) around a pool of live photons held as a structure of
//                     arrays in SHARED MEMORY; CTAs claim blocks of photons from a global counter as their pools
//                     drain.  Every loop iteration of the 16 tracing warps runs two barrier-separated stages over
//                     chunks of 32 work items (stage 1's dealt round-robin, stage 2's taken from a shared counter):
//                       1. refill + classify: retired slots take the next ray from a shared-memory ring of fresh
//                          rays; live photons draw the step's uniforms, are intersected with every node and get
//                          their free path; the slot index is appended to the VOLUME, SURFACE or EXIT queue (warp
//                          ballots + one or two shared-memory atomics per warp);
//                       2. interact: the queues are cut into chunks of 32 entries, so every warp executes ONE kind
//                          of interaction with all lanes busy, instead of all kinds with a few lanes each; an
//                          event to be tallied only writes a 64-byte request.
//                     A fifth warpgroup (service warps) tallies the requests an iteration later and keeps the ray
//                     ring filled (array loads or on-device emission).  The last photons of a CTA are drained
//                     lane by lane without barriers.  The scene blob is staged into shared memory once per CTA by a
//                     single TMA bulk copy (cp.async.bulk + mbarrier).  Random numbers are addressed by (photon,
//                     step, purpose), so neither the regrouping nor the CTA count is visible in the results.
//   trace_kernel      the same physics with one photon per lane held in registers (persistent threads, warp
//                     ballot refill from a global counter).  Handles everything the pool kernel does not:
//                     the reference's sequential xoshiro stream, more than 64 recorders, scenes too large for
//                     shared memory.  Also the cross-check of the wavefront kernel in the tests.
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
constexpr uint32_t kMaxIdleIterations = 4000000u;  // ~ seconds of polling for host rays before a CTA gives up

struct TraceArgs {
  Header hdr;          // copy of the blob's header (read as kernel-parameter constants)
  const double* blob;  // scene blob in global memory
  int blob_words;
  int scene_in_smem;   // 0: the blob did not fit into shared memory, read it through L1/L2 instead
  // Streaming upload (wavefront kernel only): the kernel is launched BEFORE the host rays have arrived.  The copy
  // engine delivers the arrays front to back, chunk by chunk, and after each chunk stores -- in stream order -- how
  // many leading rays are complete into *arrived.  CTAs claim blocks of photons in index order from `work_counter`
  // and never touch a ray at or beyond the mark.  arrived == nullptr: all rays are there.
  const uint32_t* arrived;
  const double* pos;   // [n,3]; all three null => rays are sampled on the device
  const double* dir;
  const double* wl;
  // Columns of the caller's arrays that hold ONE value for every ray (a point source, a monochromatic light) never
  // cross PCIe: the host passes the value instead (bit 0: positions, 1: directions, 2: wavelengths; the matching
  // pointer is then unused).  pvt_trace_bundle finds them (pvt_api.cu, "constant columns").
  uint32_t const_mask;
  double cpos[3], cdir[3], cwl;
  long long n, first_index, record_every;
  RunSeed keys;  // the run's seed and its window of the Philox counter space
  StepParams sp;
  u64* work_counter;
  u64* slabs;       // [gridDim.x][10 R] CTA-private tally slabs, zero on entry
  double* requests; // [gridDim.x][kReqWords][pool slots] tally requests of one iteration (service-warp kernels)
  u64* g_distinct;  // [R]
  u64* g_cross;     // [R]
  double* g_sums;   // [R,8]
  u64* g_bins;      // [total_bins]
  u64* g_stats;     // [PVT_NSTATS]
  LogColumns log;
  __host__ __device__ __forceinline__ bool has_rays() const { return pos || dir || wl || const_mask == 7u; }
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

__device__ __forceinline__ TallySink cta_sink(const TraceArgs& a, int R) {
  u64* slab = a.slabs + (size_t)blockIdx.x * 10 * R;
  return TallySink{slab, slab + R, reinterpret_cast<double*>(slab + 2 * R), a.g_bins};
}

// adds the CTA's slab into the context accumulators and the lanes' statistics into g_stats
__device__ __forceinline__ void retire_cta(const TraceArgs& a, int R, const LaneStats& st) {
  u64 s_steps = st.steps, s_events = st.events, s_rays = st.rays;  // (zero for threads that counted in shared memory)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s_steps += __shfl_down_sync(kFullMask, s_steps, off);
    s_events += __shfl_down_sync(kFullMask, s_events, off);
    s_rays += __shfl_down_sync(kFullMask, s_rays, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(a.g_stats + PVT_STAT_STEPS, s_steps);
    atomicAdd(a.g_stats + PVT_STAT_EVENTS, s_events);
    atomicAdd(a.g_stats + PVT_STAT_RAYS, s_rays);
  }
  __threadfence();  // this thread's reductions into the slab are performed before the barrier below
  __syncthreads();
  const u64* slab = a.slabs + (size_t)blockIdx.x * 10 * R;
  for (int k = threadIdx.x; k < R * 10; k += blockDim.x) {
    const u64 v = __ldcg(slab + k);
    if (k < R) { if (v) atomicAdd(a.g_distinct + k, v); }
    else if (k < 2 * R) { if (v) atomicAdd(a.g_cross + (k - R), v); }
    else {
      const double x = __longlong_as_double((long long)v);
      if (x != 0.0) atomicAdd(a.g_sums + (k - 2 * R), x);
    }
  }
}

// Is ray i of the bundle sampled for the event log, and which recorded ray is it?  (64-bit division: out of line,
// once per photon.)
static __device__ __noinline__ int sampled_ordinal(long long i, long long record_every) {
  if (record_every <= 0 || i % record_every != 0) return -1;
  return (int)(i / record_every);
}

// =========================================================================================================
// wavefront_kernel

// Shared-memory pool of P photon slots, structure of arrays: 12 f64 columns, the seen mask, six 32-bit columns,
// three u16 queues; then the ring of prefetched initial rays (7 f64 columns of K entries) and the counters.
constexpr int kPoolDoubles = 12;  // px py pz dx dy dz wl travelled duration | plan of the step: t, u, alpha
constexpr int kPoolWords = 6;     // count (< 0: slot is empty), source, nlog, idx, ids, log_ray (< 0: not sampled)
#ifndef PVT_RING
#define PVT_RING 512  // (256: +13 % on config 2 -- refills outrun the ring and fetch their rays themselves)
#endif
__host__ __device__ constexpr int ring_size(int P) { return (P > 512 && P <= 1152) ? PVT_RING : 256; }
// Shared memory left over is L1: the kernel is sensitive to it (24 KB more of shared memory cost 5 %), so the pool
// carries nothing it does not need.
constexpr int kStatWordsPerThread = 3;  // steps, events, rays of every tracing thread (SmemStats)
__host__ __device__ constexpr size_t pool_bytes(int P, int T) {
  return ((size_t)P * (kPoolDoubles * 8 + 8 /*seen*/ + kPoolWords * 4 + 3 * 2 /*queues*/ + 1 /*tallied*/) +
          (size_t)ring_size(P) * 7 * 8 + 128 /*counters*/ + 128 /*decoy words of atoms_add_by*/ +
          (size_t)kStatWordsPerThread * 4 * T + 127) / 128 * 128;
}
// the pool follows the blob on a 128-byte line, whatever the size of the scene
__host__ __device__ inline size_t pool_offset(int blob_words) { return (16 + (size_t)blob_words * 8 + 127) / 128 * 128; }
__host__ __device__ inline size_t wavefront_smem_bytes(int blob_words, int P, int T) {
  return pool_offset(blob_words) + pool_bytes(P, T);
}

// counters (u32): [0..1], [4..5] queue lengths (VOLUME | SURFACE << 16, EXIT), double buffered by iteration parity; then
// the cursors of the CTA's ray SEQUENCE.  A CTA does not own a fixed share of the bundle: it claims blocks of kClaim
// consecutive photons from a global counter as its pool drains (so all CTAs run dry together, whatever their
// photons turned out to cost) and numbers the rays it has claimed 0, 1, 2, ...; block j of that sequence starts at
// photon counters[kCtrBlock + (j & 3)].
constexpr uint32_t kClaim = 512;
enum { kCtrNext = 8, kCtrNextSnap = 9, kCtrRingHi = 10, kCtrRingHiPending = 11, kCtrSteal = 12 /* [2]: stage 1, 2 */,
       kCtrAvail = 14 /* sequence entries claimed AND arrived (snapshot taken at the last barrier) */,
       kCtrClaimed = 15, kCtrExhausted = 16 /* the global counter has run past n */, kCtrBlock = 17 /* [4] */,
       // service warps (tally requests): batch id published, requests in it, chunk cursor, chunks done, warps that
       // have left the batch, exit flag
       kSvcBatch = 21, kSvcCount = 22, kSvcSteal = 23, kSvcDone = 24, kSvcAck = 25, kSvcExit = 26,
       kSvcRayLo = 27, kSvcRayHi = 28 /* sequence entries the service warps are asked to put into the ring */, kCtrCount = 29 };
static_assert(kCtrCount <= 32, "the counters live in 128 bytes");

// photon index (within the bundle) of entry q of the CTA's sequence
__device__ __forceinline__ uint32_t sequence_photon(const uint32_t* counters, uint32_t q) {
  return counters[kCtrBlock + ((q / kClaim) & 3u)] + (q % kClaim);
}

// One thread, between stages: claim another block when the look-ahead runs short, then advance `avail` over the
// claimed entries whose photons have arrived.  `look_ahead` = ring size: production runs that far ahead of refills.
__device__ __forceinline__ void extend_sequence(const TraceArgs& a, uint32_t* counters, uint32_t look_ahead);


struct PoolView {
  double *px, *py, *pz, *dx, *dy, *dz, *wl, *trav, *dur, *t, *u, *alpha;
  u64* seen;
  int32_t *count, *source, *nlog, *log_ray;
  uint32_t *idx, *ids;
  uint16_t *qv, *qs, *qe;
  uint8_t* tallied;    // service-warp kernels: this photon has sent a tally request before (its seen mask is live)
  double* ring;        // [7][K]: px py pz dx dy dz wl of entries [.., ring_hi) of the CTA's ray sequence, at entry mod K
  uint32_t* counters;
  uint32_t* stats;     // [3][T] steps, events, rays per tracing thread (SmemStats), behind the counters and decoys
};

__device__ __forceinline__ PoolView carve_pool(unsigned char* base, int P) {
  PoolView v;
  double* d = reinterpret_cast<double*>(base);
  v.px = d; v.py = d + P; v.pz = d + 2 * P; v.dx = d + 3 * P; v.dy = d + 4 * P; v.dz = d + 5 * P;
  v.wl = d + 6 * P; v.trav = d + 7 * P; v.dur = d + 8 * P; v.t = d + 9 * P; v.u = d + 10 * P; v.alpha = d + 11 * P;
  v.seen = reinterpret_cast<u64*>(d + 12 * P);
  v.ring = d + 13 * P;
  int32_t* w = reinterpret_cast<int32_t*>(v.ring + 7 * ring_size(P));
  v.count = w; v.source = w + P; v.nlog = w + 2 * P;
  v.idx = reinterpret_cast<uint32_t*>(w + 3 * P); v.ids = reinterpret_cast<uint32_t*>(w + 4 * P);
  v.log_ray = w + 5 * P;
  uint16_t* h = reinterpret_cast<uint16_t*>(w + 6 * P);
  v.qv = h; v.qs = h + P; v.qe = h + 2 * P;
  v.tallied = reinterpret_cast<uint8_t*>(h + 3 * P);
  v.counters = reinterpret_cast<uint32_t*>(v.tallied + P);  // P is a multiple of 32: 4-byte aligned
  v.stats = v.counters + 64;
  return v;
}

// A pool photon's path length and time of flight stay in their pool columns (MemAcc): -2 % on config 2, -1.2 % on
// the validation scene against loading them into registers for the length of the interaction.
typedef PhotonT<2, MemAcc> PoolPhoton;
constexpr bool kAccInPool = true;
// binds the accumulators of a pool photon to its slot (before anything reads or writes them)
__device__ __forceinline__ void bind_slot(const PoolView& pool, int s, PhotonT<2, MemAcc>& ph) {
  ph.travelled.at = pool.trav + s; ph.duration.at = pool.dur + s;
}
__device__ __forceinline__ void bind_slot(const PoolView&, int, PhotonT<2>&) {}

// what classify needs of a slot
template <bool kLog>
__device__ __forceinline__ void load_slot_head(const PoolView& pool, int s, PoolPhoton& ph, int max_events) {
  ph.p = V3{pool.px[s], pool.py[s], pool.pz[s]};
  ph.d = V3{pool.dx[s], pool.dy[s], pool.dz[s]};
  ph.wl = pool.wl[s];
  ph.count = pool.count[s];
  ph.log_ray = -1; ph.log_base = -1; ph.nlog = 0;
  bind_slot(pool, s, ph);
  if (kLog) {
    if (!kAccInPool) { ph.travelled = pool.trav[s]; ph.duration = pool.dur[s]; }
    ph.source = pool.source[s];
    ph.nlog = pool.nlog[s];
    ph.log_ray = pool.log_ray[s];
    ph.log_base = ph.log_ray < 0 ? -1 : (long long)ph.log_ray * max_events;
  }
}
template <bool kLog, bool kSeen = true>
__device__ __forceinline__ void load_slot(const PoolView& pool, int s, PoolPhoton& ph, int max_events) {
  load_slot_head<kLog>(pool, s, ph, max_events);
  if (!kAccInPool) { ph.travelled = pool.trav[s]; ph.duration = pool.dur[s]; }
  ph.source = pool.source[s];
  const u64 seen = kSeen ? pool.seen[s] : 0ull;  // (the service warps own the masks when there are any)
  ph.seen[0] = (uint32_t)seen; ph.seen[1] = (uint32_t)(seen >> 32);
}
template <bool kLog>
__device__ __forceinline__ void store_slot(const PoolView& pool, int s, const PoolPhoton& ph) {
  pool.px[s] = ph.p.x; pool.py[s] = ph.p.y; pool.pz[s] = ph.p.z;
  pool.dx[s] = ph.d.x; pool.dy[s] = ph.d.y; pool.dz[s] = ph.d.z;
  pool.wl[s] = ph.wl;
  if (!kAccInPool) { pool.trav[s] = ph.travelled; pool.dur[s] = ph.duration; }
  pool.source[s] = ph.source;
  if (kLog) pool.nlog[s] = ph.nlog;
}

// Shared-memory atomics as single instructions.  (atomicAdd() by one elected lane makes the compiler wrap its own
// warp aggregation -- vote, elect, popc, shuffle, a dozen one-lane instructions -- around every call.)
__device__ __forceinline__ uint32_t atoms_add(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_addr(p)), "r"(v) : "memory");
  return old;
}

// One elected lane adds `v` to a shared counter and every lane gets the old value -- executed by ALL 32 lanes: the
// others add zero to a decoy word of their own (counters + 32 + lane).  ptxas wraps its own warp aggregation (vote,
// find-leader, popc, predicate, shuffle, add: a dozen dependent instructions) around any shared add whose address is
// the same in every lane, inline PTX included; with an address that differs per lane it emits the single ATOMS.
__device__ __forceinline__ uint32_t atoms_add_by(int leader, int lane, uint32_t* counters, int word, uint32_t v) {
  const bool mine = lane == leader;
  const uint32_t old = atoms_add(counters + (mine ? word : 32 + lane), mine ? v : 0u);
  return __shfl_sync(kFullMask, old, leader);
}

// The VOLUME and SURFACE queue lengths share one 32-bit word (VOLUME | SURFACE << 16; the EXIT queue has its own), so
// a classified chunk appends to the queues with one or two independent shared atomics.  (A 64-bit word for all three
// would be a compare-and-swap loop: there is no native 64-bit shared-memory add.)
__device__ __forceinline__ void push_queues(const uint16_t* qv_, const uint16_t* qs_, const uint16_t* qe_, uint32_t* counters,
                                            int word, int cls, int s, int lane) {
  uint16_t* qv = const_cast<uint16_t*>(qv_); uint16_t* qs = const_cast<uint16_t*>(qs_); uint16_t* qe = const_cast<uint16_t*>(qe_);
  const unsigned mv = __ballot_sync(kFullMask, cls == kVolume), ms = __ballot_sync(kFullMask, cls == kSurface),
                 me = __ballot_sync(kFullMask, cls == kExit || cls == kKill);
  uint32_t base_vs = 0, base_e = 0;
  if (mv | ms) base_vs = atoms_add_by(0, lane, counters, word, (uint32_t)__popc(mv) | ((uint32_t)__popc(ms) << 16));
  if (me) base_e = atoms_add_by(0, lane, counters, word + 1, (uint32_t)__popc(me));
  const unsigned below = (1u << lane) - 1u;
  // One predicated store through selects.  Written as three `if (cls == ...) queue[...] = s` the compiler builds a jump
  // table on `cls` (BRX: an indirect, divergent branch in every classified chunk); this form measured -5.2 % on config 2.
  const bool is_v = cls == kVolume, is_s = cls == kSurface;
  const unsigned mine = is_v ? mv : (is_s ? ms : me);
  uint16_t* q = is_v ? qv : (is_s ? qs : qe);
  const uint32_t base = is_v ? (base_vs & 0xffffu) : (is_s ? (base_vs >> 16) : base_e);
  if (cls != kDead) q[base + __popc(mine & below)] = (uint16_t)s;
}

// initial state of photon i of the bundle (global arrays or the emitter); out of line: the common path takes
// fresh rays from the shared-memory ring that the spare warps keep filled
static __device__ __noinline__ void fetch_ray(const TraceArgs& a, const SceneView sv, long long i, V3& p, V3& d, double& wl) {
  if (a.has_rays()) {
    p = (a.const_mask & 1u) ? V3{a.cpos[0], a.cpos[1], a.cpos[2]}
                            : V3{__ldcg(a.pos + 3 * i), __ldcg(a.pos + 3 * i + 1), __ldcg(a.pos + 3 * i + 2)};
    d = (a.const_mask & 2u) ? V3{a.cdir[0], a.cdir[1], a.cdir[2]}
                            : V3{__ldcg(a.dir + 3 * i), __ldcg(a.dir + 3 * i + 1), __ldcg(a.dir + 3 * i + 2)};
    wl = (a.const_mask & 4u) ? a.cwl : __ldcg(a.wl + i);
  } else {
    emit_ray(sv, a.keys, a.first_index + i, p, d, wl);
  }
}

__device__ __forceinline__ void extend_sequence(const TraceArgs& a, uint32_t* counters, uint32_t look_ahead) {
  const uint32_t next = counters[kCtrNext];
  uint32_t claimed = counters[kCtrClaimed];
  // a new block J overwrites the table entry of block J - 4, whose entries must all have been handed out:
  // claimed - next <= 3 kClaim.  (Refills happen in stage 1 only; this runs in stage 2.)
  if (!counters[kCtrExhausted] && claimed - next < look_ahead + kClaim && claimed - next <= 3u * kClaim) {
    const u64 b = atomicAdd(a.work_counter, (u64)kClaim);
    if (b >= (u64)a.n) {
      counters[kCtrExhausted] = 1u;
    } else {
      counters[kCtrBlock + ((claimed / kClaim) & 3u)] = (uint32_t)b;
      const u64 left = (u64)a.n - b;
      claimed += left < kClaim ? (uint32_t)left : kClaim;
      if (left <= kClaim) counters[kCtrExhausted] = 1u;  // that was the last (possibly partial) block
      counters[kCtrClaimed] = claimed;
    }
  }
  uint32_t avail = counters[kCtrAvail];
  if (!a.arrived) {
    avail = claimed;
  } else {
    const uint32_t mark = *reinterpret_cast<const volatile uint32_t*>(a.arrived);
    while (avail < claimed) {  // at most four blocks
      const uint32_t b = counters[kCtrBlock + ((avail / kClaim) & 3u)], block_lo = avail - avail % kClaim;
      const uint32_t block_n = claimed - block_lo < kClaim ? claimed - block_lo : kClaim;
      const uint32_t there = mark > b ? (mark - b < block_n ? mark - b : block_n) : 0u;
      if (block_lo + there <= avail) break;
      avail = block_lo + there;
      if (there < block_n) break;
    }
  }
  counters[kCtrAvail] = avail;
}

// take the next chunk of 32 work items of the current stage: one shared atomic per warp
__device__ __forceinline__ uint32_t steal_chunk(uint32_t* counters, int word, int lane) {
  return atoms_add_by(0, lane, counters, word, 1u);
}

// ---- service warps --------------------------------------------------------------------------------------------
// Tallying in place costs a stage-2 chunk a third of its latency for the few lanes that have something to tally.
// Kernels instantiated with S > 0 run S extra threads (one warpgroup) that do nothing else: stage 2 only WRITES a
// request (64 bytes: wavelength, duration, travelled, cosine, local point, packed ids) into a per-CTA ring in global
// memory (L2, two halves by iteration parity); the barrier that ends the stage makes the batch visible, thread 0
// publishes it at the top of the next iteration, and the service warps tally it 32 requests at a time while the
// tracing warps go on.  They have a whole iteration to do so (the half is rewritten two stages later), so nobody
// waits for anybody in the steady state.  The distinct-ray (`seen`) masks are theirs alone: a request only says
// whether it is the photon's first, batches are served in order and a slot appears once per batch.  The tracing
// warps synchronise among themselves on a named barrier, and the registers of the CTA are re-divided between the
// two roles (setmaxnreg).
#ifndef PVT_DRAIN_PHOTONS
#define PVT_DRAIN_PHOTONS 32  // (32 / 96 / 256 / 512 measured 6.45 / 6.51 / 6.51 / 6.56 ms on config 2)
#endif
constexpr int kDrainPhotons = PVT_DRAIN_PHOTONS;  // a CTA with no supply left and at most this many live photons drains them lane by lane
constexpr int kReqWords = 8;
// Registers after the re-division (the pool is per CTA: the two sides must add up to what the launch allocated):
//   S = 128: launched as 640 x 96, becomes 512 x 112 + 128 x 32 for scenes of axis-aligned boxes (every LSC), 512 x 104 +
//   128 x 64 otherwise.  The tracing warps set the pace and every register they lack is a spill on their critical path;
//   in the LSC scenes the service warps have slack (0.25 tally requests per photon step: they sleep on their barrier a
//   tenth of the time) and absorb the spills instead: 104 + 64 -> 112 + 32 measured -6 to -8 % (config 2: 6.69 -> 6.16 ms,
//   validation 17.5 -> 16.4 ms, within one lease).  Where nearly every step tallies (hello_world: one request per photon
//   step) the service side is the busier one and the same move costs 15 % (4.84 -> 5.56 ms).
// setmaxnreg is a WARPGROUP instruction (four warps execute it together): a service side of two warps (S = 64, launched
// as 576 x 112) cannot re-divide -- it hangs -- and keeps the uniform 112.
__host__ __device__ constexpr bool redivide_regs(int S) { return S > 0 && S % 128 == 0; }
__host__ __device__ constexpr int tracer_regs(bool boxes) { return boxes ? 112 : 104; }
__host__ __device__ constexpr int service_regs(bool boxes) { return boxes ? 32 : 64; }

#ifndef PVT_BOX_SURFACE
#define PVT_BOX_SURFACE 1
#endif
constexpr bool kBoxSurface = PVT_BOX_SURFACE != 0;  // surface_step / exit_step specialised for axis-aligned boxes

template <int T, int S>
__device__ __forceinline__ void sync_tracers() {
  if (S == 0) __syncthreads();
  else asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
}
template <int T, int S>
__device__ __forceinline__ int sync_tracers_count(bool pred) {
  if (S == 0) return __syncthreads_count(pred);
  int total;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.popc.u32 %0, 1, %2, q;\n\t}"
               : "=r"(total) : "r"((int)pred), "n"(T) : "memory");
  return total;
}
__device__ __forceinline__ uint32_t peek(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// what stage 2 leaves for the service warps
__device__ __forceinline__ void write_request(double* ring, int stride, uint32_t at, const PoolPhoton& ph, const TallyReq& tr,
                                              int slot, bool fresh) {
  double* q = ring + at;
  const u64 packed = (u64)(uint32_t)tr.sel | ((u64)(uint32_t)tr.node << 8) | ((u64)(uint32_t)(tr.face + 1) << 16) |
                     ((u64)(tr.has_normal ? 1u : 0u) << 20) | ((u64)(fresh ? 1u : 0u) << 21) | ((u64)(uint32_t)slot << 32);
  q[0] = ph.wl; q[stride] = ph.duration; q[2 * stride] = ph.travelled; q[3 * stride] = tr.cosine;
  q[4 * stride] = tr.lp.x; q[5 * stride] = tr.lp.y; q[6 * stride] = tr.lp.z;
  q[7 * stride] = __longlong_as_double((long long)packed);
}

// entry o of the CTA's sequence into the ring of fresh rays (K entries, seven columns)
// kInlineEmitter: the service warps inline the emitter (no out-of-line call inside their reduced register allocation;
// the emitter itself takes sin(cone angle) from the light record -- sin()'s slow path inside a setmaxnreg region
// crashes ptxas 12.9)
template <int K, bool kInlineEmitter = false>
__device__ __forceinline__ void produce_ray(const TraceArgs& a, const SceneView& sv, const PoolView& pool, uint32_t o) {
  const long long i = sequence_photon(pool.counters, o);
  double* r = pool.ring + (o & (K - 1));
  if (a.has_rays()) {
    // one ray per lane.  Word-granular cooperative loads of the chunk's columns (lane l takes words l, l + 32, l + 64 of the
    // contiguous run and scatters them into the ring: every sector fetched once instead of three times) were measured
    // twice and lost twice -- round 1, and round 2 in the service warps: +2 % on config 2, +13 % on hello_world, where
    // the service side is the busy one.  The L2 sectors re-touched cost nothing; the index arithmetic does.
    // L2-only loads: with a streaming upload a cached line could hold a neighbour that had not arrived
    if (a.const_mask & 1u) { r[0] = a.cpos[0]; r[K] = a.cpos[1]; r[2 * K] = a.cpos[2]; }
    else { r[0] = __ldcg(a.pos + 3 * i); r[K] = __ldcg(a.pos + 3 * i + 1); r[2 * K] = __ldcg(a.pos + 3 * i + 2); }
    if (a.const_mask & 2u) { r[3 * K] = a.cdir[0]; r[4 * K] = a.cdir[1]; r[5 * K] = a.cdir[2]; }
    else { r[3 * K] = __ldcg(a.dir + 3 * i); r[4 * K] = __ldcg(a.dir + 3 * i + 1); r[5 * K] = __ldcg(a.dir + 3 * i + 2); }
    r[6 * K] = (a.const_mask & 4u) ? a.cwl : __ldcg(a.wl + i);
  } else if (kInlineEmitter) {
    const EmittedRay e = emit_ray_value(sv, a.keys, a.first_index + i);
    r[0] = e.pos.x; r[K] = e.pos.y; r[2 * K] = e.pos.z;
    r[3 * K] = e.dir.x; r[4 * K] = e.dir.y; r[5 * K] = e.dir.z;
    r[6 * K] = e.wl;
  } else {
    emit_ray_to_ring(sv, a.keys, a.first_index + i, r, K);
  }
}

template <int P, int K, int kSvcWarps>
__device__ __forceinline__ void service_loop(const TraceArgs& a, const SceneView& sv, const TallySink& sink,
                                             const PoolView& pool, const double* ring, int lane) {
  for (;;) {
    // sleep on named barrier 2 until warp 0 of the tracers arrives with a batch (or with the exit flag): no polling
    asm volatile("bar.sync 2, %0;" ::"n"(kSvcWarps * 32 + 32) : "memory");
    if (peek(pool.counters + kSvcExit)) return;
    const uint32_t batch = peek(pool.counters + kSvcBatch);
    const uint32_t count = peek(pool.counters + kSvcCount), chunks = (count + 31u) >> 5;
    const uint32_t ray_lo = peek(pool.counters + kSvcRayLo), ray_hi = peek(pool.counters + kSvcRayHi);
    const uint32_t ray_chunks = (ray_hi - ray_lo + 31u) >> 5;
    const double* half = ring + (size_t)(batch & 1u) * kReqWords * P;  // batch b = requests of iteration b - 2
    for (;;) {
      uint32_t chunk = steal_chunk(pool.counters, kSvcSteal, lane);
      if (chunk >= ray_chunks + chunks) break;
      if (chunk < ray_chunks) {  // fresh rays first: the tracing warps refill from them next iteration
        const uint32_t o = ray_lo + chunk * 32u + (uint32_t)lane;
        if (o < ray_hi) produce_ray<K, true>(a, sv, pool, o);
        __syncwarp();
        continue;
      }
      chunk -= ray_chunks;
      const uint32_t at = chunk * 32u + (uint32_t)lane;
      if (at < count) {
        const double* q = half + at;
        const double wl = __ldcg(q), duration = __ldcg(q + P), travelled = __ldcg(q + 2 * P), cosine = __ldcg(q + 3 * P);
        const V3 lp = V3{__ldcg(q + 4 * P), __ldcg(q + 5 * P), __ldcg(q + 6 * P)};
        const u64 packed = (u64)__double_as_longlong(__ldcg(q + 7 * P));
        const int sel = (int)(packed & 0xffu), node = (int)((packed >> 8) & 0xffu), face = (int)((packed >> 16) & 0xfu) - 1;
        const bool has_normal = (packed >> 20) & 1u, fresh = (packed >> 21) & 1u;
        const int slot = (int)(packed >> 32);
        // the seen masks belong to the service warps: batches are served in order and a slot appears once per batch
        const u64 seen_bits = fresh ? 0ull : pool.seen[slot];
        V3 nw = V3{0.0, 0.0, 0.0};
        if (has_normal && face < 0) {  // not a box face: the facet test needs the world normal, recomputed from the point
          const double* rec = sv.node(node);
          nw = map_vector(rec + kNodeL2W, outward_normal(sv.node_int(node, NI_GEOM), rec + kNodeParams, lp));
        }
        SeenMask<2> seen;
        seen.w[0] = (uint32_t)seen_bits; seen.w[1] = (uint32_t)(seen_bits >> 32);
        seen = tally_event<2>(sv, sink, seen, sel, node, face, has_normal, nw, lp, cosine, wl, duration, travelled);
        pool.seen[slot] = (u64)seen.w[0] | ((u64)seen.w[1] << 32);
      }
      __syncwarp();
    }
    __threadfence_block();
    if (lane == 0) atoms_add(pool.counters + kSvcAck, 1u);
  }
}

// T tracing threads per CTA, P pool slots (multiple of 32, typically ~2T), B resident CTAs per SM, S service threads
template <int T, int P, int B, bool kLog, bool kBoxes = false, int S = 0>
__global__ void __launch_bounds__(T + S, B) wavefront_kernel(const __grid_constant__ TraceArgs a) {
  constexpr int K = ring_size(P);
  constexpr int kSvcWarps = S / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [mbarrier 16 B][blob][pool, on the next 128-byte line].  (The other order -- pool first, so that its columns and
  // the blob sit at compile-time addresses -- was measured: ptxas spills more with the immediates, +13 %.)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sblob = reinterpret_cast<double*>(smem_raw + 16);
  stage_blob(sblob, a.blob, (uint32_t)a.blob_words * 8u, bar);
  const SceneView sv{sblob, &a.hdr};
  const PoolView pool = carve_pool(smem_raw + pool_offset(a.blob_words), P);
  const int R = sv.hdr().n_recorders;
  const TallySink sink = cta_sink(a, R);
  const u64 id0 = (u64)a.first_index;  // photon index of ray 0 of the bundle within the run
  const StepParams sp = a.sp;
  const LogColumns& L = a.log;

  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < kCtrCount) pool.counters[tid] = tid == kSvcAck ? (uint32_t)kSvcWarps : 0u;
  for (int s = tid; s < P; s += T + S) pool.count[s] = -1;
  for (int k = tid; k < kStatWordsPerThread * T; k += T + S) pool.stats[k] = 0u;
  __syncthreads();
  if (tid == 0) {  // the first blocks of this CTA's sequence: enough to fill the pool
    for (int k = 0; k < (P + (int)kClaim - 1) / (int)kClaim && k < 3; ++k) extend_sequence(a, pool.counters, (uint32_t)P);
  }
  __syncthreads();

  // run statistics in shared memory, not in registers (see SmemStats; with early_u: -1.2 % config 2, -4.5 % validation)
  SmemStats<T> st{smem_addr(pool.stats) + 4u * (uint32_t)tid};  // (service threads, tid >= T, never count)
  double* const ring = S > 0 ? a.requests + (size_t)blockIdx.x * 2 * kReqWords * P : nullptr;  // two halves, by parity
  const bool service = S > 0 && tid >= T;
  const bool svc_rays = S > 0;  // the service warps also fill the ring of fresh rays
  if (service) {
    if (redivide_regs(S)) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(service_regs(kBoxes)));
    service_loop<P, K, kSvcWarps>(a, sv, sink, pool, ring, lane);
    __threadfence();
  } else {  // the tracing warps; both roles meet again at retire_cta's barrier below
  if (redivide_regs(S)) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(tracer_regs(kBoxes)));
  uint32_t idle_iterations = 0;
  bool draining = false;
#ifdef PVT_PROFILE_STAGES  // where warp 0's time goes: stage 1, its barrier, stage 2, its barrier (cycles) -> stats[4..7]
  long long prof[4] = {0, 0, 0, 0}, prof_t = clock64();
  u64 prof_start;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_start));
#define PVT_PROF(k) do { const long long now_ = clock64(); prof[k] += now_ - prof_t; prof_t = now_; } while (0)
#else
#define PVT_PROF(k) do { } while (0)
#endif
  for (uint32_t iter = 0;; ++iter) {
    uint32_t* qn = pool.counters + 4 * (iter & 1);  // queue lengths of this iteration: VOLUME | SURFACE << 16, EXIT
    // ---------------- stage 1: refill + classify the pool, 32 slots per chunk ----------------------------------
    // (Kernels without service warps also produce the fresh rays here, first, and take both kinds of chunk from one
    // shared counter; with service warps the classification chunks are dealt round-robin, see below.)
    bool live = false;
    if (S > 0 && tid < 32) {
      // Warp 0 hands the service warps their work for this iteration and wakes them: the tally requests written by
      // the stage 2 that just ended (its barrier made them visible; ring half by parity) and the next stretch of
      // fresh rays to put into the ring.  They have the whole iteration for it, so the wait for the PREVIOUS
      // hand-over's acknowledgements almost never spins; once it is in, what was asked for then is in the ring.
      uint32_t work = 0;
      if (tid == 0) {
        while (peek(pool.counters + kSvcAck) < (uint32_t)kSvcWarps) { }
        uint32_t lo = 0, hi = 0;
        if (svc_rays) {
          const uint32_t produced = pool.counters[kSvcRayHi], next = pool.counters[kCtrNextSnap];
          // (the other warps may still read the old value this iteration: they then fetch those rays themselves)
          pool.counters[kCtrRingHi] = produced;
          // to produce: [max(produced, next), next + K), none beyond what has arrived; `next` is the snapshot taken
          // at the last barrier, so ring entries that this iteration's refills read are never overwritten
          lo = produced > next ? produced : next;
          const uint32_t avail = pool.counters[kCtrAvail];
          hi = next + (uint32_t)K;
          if (hi > avail) hi = avail;
          if (hi < lo) hi = lo;
          pool.counters[kSvcRayLo] = lo;
          pool.counters[kSvcRayHi] = hi;
        }
        const uint32_t count = pool.counters[4 * ((iter + 1) & 1) + 2];
        work = count + (hi - lo);
        if (work) {  // nothing to hand over (a draining or starving CTA): the service warps sleep on
          pool.counters[kSvcAck] = 0u;
          pool.counters[kSvcCount] = count;
          pool.counters[kSvcSteal] = 0u;
          pool.counters[kSvcBatch] = iter + 1u;
          __threadfence_block();
        }
      }
      work = __shfl_sync(kFullMask, work, 0);
      if (work) asm volatile("bar.arrive 2, %0;" ::"n"(kSvcWarps * 32 + 32) : "memory");
    }
#ifndef PVT_NO_DRAIN
    if (iter > 0) {
      // The last photons of a CTA: no ray will ever be refilled and few slots are live (the previous iteration's
      // queue lengths, still in place, bound their number).  Two barriers and two rounds of work stealing per step
      // are then pure latency in front of a handful of serial chains: leave the loop and let every thread run the
      // photons of its own slots to their ends (below).
      const uint32_t* prev = pool.counters + 4 * ((iter + 1) & 1);
      const uint32_t live_before = (prev[0] & 0xffffu) + (prev[0] >> 16) + prev[1];
      if (live_before <= (uint32_t)kDrainPhotons && pool.counters[kCtrExhausted] &&
          pool.counters[kCtrNextSnap] >= pool.counters[kCtrClaimed]) {
        draining = true;
        break;
      }
    }
#endif
    {
      const uint32_t ring_hi = pool.counters[kCtrRingHi];  // entries [.., ring_hi) of the sequence are in the ring
      // rays to produce: [max(ring_hi, next), next + K), `next` being the snapshot taken at the last barrier, so
      // ring entries read by this stage's refills are never overwritten
      const uint32_t next = pool.counters[kCtrNextSnap];
      const uint32_t lo = ring_hi > next ? ring_hi : next;
      const uint32_t avail = pool.counters[kCtrAvail];  // claimed and arrived
      uint32_t hi = next + (uint32_t)K;
      if (hi > avail) hi = avail;
      if (hi < lo) hi = lo;
      const uint32_t classify_chunks = P / 32, ray_chunks = svc_rays ? 0u : (hi - lo + 31u) >> 5;
      if (tid == 0) {
        if (!svc_rays) pool.counters[kCtrRingHiPending] = hi;
        pool.counters[kCtrSteal + 1] = 0u;  // stage 2's work counter
      }
      // Kernels whose service warps fill the ring: the stage is the P / 32 classification chunks alone, all of one
      // cost -- a fixed round-robin share per warp, no counter, no atomic, no failing last steal (-2 % on config 2
      // against taking them from the shared counter like the chunks of stage 2, whose costs differ by kind).
      uint32_t static_chunk = (uint32_t)(tid >> 5);
      for (;;) {
        uint32_t chunk;
        if (svc_rays) { chunk = static_chunk; static_chunk += (uint32_t)(T / 32); }
        else chunk = steal_chunk(pool.counters, kCtrSteal, lane);
        if (chunk >= classify_chunks + ray_chunks) break;
        // Ray production first: its loads have the longest latency of the stage (HBM, or host memory over PCIe
        // when the caller's arrays are page-locked) and the warp that waits for them simply takes fewer
        // classification chunks afterwards (10.33 -> 10.19 ms on config 2 against producing last).
        if (chunk < ray_chunks) {
          const uint32_t o = lo + chunk * 32u + (uint32_t)lane;  // sequence entry of this lane's ray
          if (o < hi) produce_ray<K>(a, sv, pool, o);
          continue;
        }
        chunk -= ray_chunks;
        if (chunk < classify_chunks) {
          const int slot = (int)(chunk * 32u) + lane;
          PoolPhoton ph;
          // (loading every lane's slot here, ahead of the refill's atomic and outside the branch below: +1 %)
          bind_slot(pool, slot, ph);
          const bool dead = pool.count[slot] < 0;
          const unsigned m = __ballot_sync(kFullMask, dead);
          bool fresh = false;
          if (m) {
            // take popc(m) ray indices, but none at or beyond `avail`: what was taken in excess is given back
            // (another warp may see the inflated cursor meanwhile and take less than it could -- its slots
            // simply stay empty until the next iteration; indices below `avail` are handed out exactly once)
            const uint32_t cnt = (uint32_t)__popc(m);
            const uint32_t base = atoms_add_by(0, lane, pool.counters, kCtrNext, cnt);
            if (lane == 0 && base + cnt > avail)
              atoms_add(pool.counters + kCtrNext, 0u - (base + cnt - (base > avail ? base : avail)));
            const uint32_t mine = base + __popc(m & ((1u << lane) - 1u));
            if (dead && mine < avail) {
              const long long i = sequence_photon(pool.counters, mine);
              if (mine < ring_hi) {
                const double* r = pool.ring + (mine & (K - 1));
                ph.p = V3{r[0], r[K], r[2 * K]};
                ph.d = V3{r[3 * K], r[4 * K], r[5 * K]};
                ph.wl = r[6 * K];
              } else {
                // (through temporaries: taking the address of ph's fields would pin the whole photon in local memory)
                V3 tp, td;
                double tw;
                fetch_ray(a, sv, i, tp, td, tw);
                ph.p = tp; ph.d = td; ph.wl = tw;
              }
              ph.log_ray = (kLog && a.record_every > 0) ? sampled_ordinal(i, a.record_every) : -1;
              ph.log_base = ph.log_ray < 0 ? -1 : (long long)ph.log_ray * sp.max_events;
              begin_photon<kLog>(ph, L, sp, st);
              pool.idx[slot] = (uint32_t)i;
              if (kLog) pool.log_ray[slot] = ph.log_ray;
              if (S > 0) pool.tallied[slot] = 0; else pool.seen[slot] = 0ull;
              fresh = true;
            }
          }
          StepClass cls = kDead;
          if (fresh || !dead) {
            if (!fresh) load_slot_head<kLog>(pool, slot, ph, sp.max_events);
            PhiloxStream rng;
            rng.init(a.keys, id0 + (u64)pool.idx[slot]);
            StepPlan plan;
            cls = classify_step<kLog, kBoxes>(sv, L, sp, ph, rng, st, plan, pool.u + slot);  // (stores the step's pool.u itself)
            if (cls == kDead) {
              if (kLog && ph.log_ray >= 0) L.counts[ph.log_ray] = ph.nlog;
              pool.count[slot] = -1;
            } else {
              if (fresh) store_slot<kLog>(pool, slot, ph);
              pool.count[slot] = ph.count;
              pool.t[slot] = plan.t; pool.alpha[slot] = plan.alpha;
              pool.ids[slot] = (uint32_t)(plan.hit & 0xff) | ((uint32_t)(plan.container & 0xff) << 8) |
                               ((uint32_t)(plan.adjacent & 0xff) << 16) | (cls == kKill ? 1u << 24 : 0u);
              live = true;
            }
          }
          push_queues(pool.qv, pool.qs, pool.qe, pool.counters, 4 * (int)(iter & 1), cls, slot, lane);
        }
      }
    }
    // leave when no slot is live and every claimed ray has been taken with nothing left to claim; while rays are still in flight from
    // the host the CTA keeps polling (bounded: a transfer that never completes must not hang the device)
    const bool pending = !(pool.counters[kCtrExhausted] && pool.counters[kCtrNextSnap] >= pool.counters[kCtrClaimed]) &&
                         idle_iterations < kMaxIdleIterations;
    PVT_PROF(0);
    const bool any_live = sync_tracers_count<T, S>(live) > 0;
    PVT_PROF(1);
    if (!any_live && !pending) break;
    idle_iterations = any_live ? 0u : idle_iterations + 1u;

    // ---------------- stage 2: interact; chunks of 32 entries of the VOLUME, SURFACE, EXIT queues in that
    // order (longest first), each chunk one kind of interaction with every lane busy ------------------------
    const uint32_t cv = qn[0] & 0xffffu, cs = qn[0] >> 16, ce = qn[1];
    const uint32_t nv = (cv + 31u) >> 5, ns = (cs + 31u) >> 5, ne = (ce + 31u) >> 5;
    if (tid < 4) pool.counters[4 * ((iter + 1) & 1) + tid] = 0u;
    if (tid == T - 1) {  // publish the cursors for the next iteration (nobody refills or produces in stage 2)
      if (!svc_rays) pool.counters[kCtrRingHi] = pool.counters[kCtrRingHiPending];
      pool.counters[kCtrNextSnap] = pool.counters[kCtrNext];
      extend_sequence(a, pool.counters, (uint32_t)K);
      pool.counters[kCtrSteal] = 0u;  // stage 1's work counter
    }
    for (;;) {  // (a fixed first chunk per warp before stealing measured +0.4 %: the kinds differ too much in cost)
      uint32_t chunk = steal_chunk(pool.counters, kCtrSteal + 1, lane);
      if (chunk >= nv + ns + ne) break;
      int slot = -1, cls;
      if (chunk < nv) {
        cls = kVolume;
        const uint32_t e = chunk * 32u + (uint32_t)lane;
        if (e < cv) slot = pool.qv[e];
      } else if (chunk < nv + ns) {
        cls = kSurface;
        const uint32_t e = (chunk - nv) * 32u + (uint32_t)lane;
        if (e < cs) slot = pool.qs[e];
      } else {
        cls = kExit;
        const uint32_t e = (chunk - nv - ns) * 32u + (uint32_t)lane;
        if (e < ce) slot = pool.qe[e];
      }
      if (slot >= 0) {
        PoolPhoton ph;
        TallyReq tr;
        bool alive = false;
        load_slot<kLog, S == 0>(pool, slot, ph, sp.max_events);
        PhiloxStream rng;
        rng.init(a.keys, id0 + (u64)pool.idx[slot]);
        rng.begin_step((uint32_t)ph.count);
        StepPlan plan;
        plan.t = pool.t[slot]; plan.u = pool.u[slot]; plan.alpha = pool.alpha[slot];
        const uint32_t ids = pool.ids[slot];
        plan.hit = (int)(ids & 0xff); plan.container = (int)((ids >> 8) & 0xff); plan.adjacent = (int)((ids >> 16) & 0xff);
        if (plan.adjacent == 0xff) plan.adjacent = -1;
        if (cls == kVolume) alive = volume_step<kLog>(sv, L, sp, ph, rng, st, plan, tr);
        else if (cls == kSurface) alive = surface_step<kLog, kBoxes && kBoxSurface>(sv, L, sp, ph, rng, st, plan, tr);
        else if (ids >> 24) kill_step<kLog>(sv, L, sp, ph, st, plan, tr);
        else exit_step<kLog, kBoxes && kBoxSurface>(sv, L, sp, ph, st, plan, tr);
        if (alive) {
          store_slot<kLog>(pool, slot, ph);
        } else {
          if (kLog && ph.log_ray >= 0) L.counts[ph.log_ray] = ph.nlog;
          pool.count[slot] = -1;
        }
        // Tallied in place (S == 0), by the few lanes of the chunk that have something to tally (~8 active lanes, a
        // fifth of the kernel's issue slots and a third of this stage's latency) -- or handed to the service warps.
        if (S == 0 && tr.sel >= 0) {
          tally(sv, sink, ph, tr);
          if (alive) pool.seen[slot] = (u64)ph.seen[0] | ((u64)ph.seen[1] << 32);
        }
        if (S > 0 && tr.sel >= 0) {
          // every requesting lane reserves its own ring entry (one shared atomic instruction for the warp, no
          // reconvergence needed) while the values are still in registers
          const uint32_t at = atoms_add(qn + 2, 1u);
          write_request(ring + (size_t)(iter & 1u) * kReqWords * P, P, at, ph, tr, slot, pool.tallied[slot] == 0);
          pool.tallied[slot] = 1;
        }
      }
    }
    PVT_PROF(2);
    sync_tracers<T, S>();
    PVT_PROF(3);
  }
#ifndef PVT_NO_DRAIN
  if (draining) {
    // ---------------- drain: one photon per lane, whole steps, no barriers, tallies in place --------------------
    if (S > 0) {  // the seen masks come back from the service warps once their last batch is done
      sync_tracers<T, S>();  // ... which warp 0 has handed over by the time it gets here
      if (lane == 0) while (peek(pool.counters + kSvcAck) < (uint32_t)kSvcWarps) { }
      __syncwarp();
      __threadfence_block();
    }
    for (int slot = tid; slot < P; slot += T) {
      if (pool.count[slot] < 0) continue;
      PoolPhoton ph;
      load_slot<kLog, S == 0>(pool, slot, ph, sp.max_events);
      if (S > 0 && pool.tallied[slot]) {
        const u64 seen = pool.seen[slot];
        ph.seen[0] = (uint32_t)seen; ph.seen[1] = (uint32_t)(seen >> 32);
      }
      PhiloxStream rng;
      rng.init(a.keys, id0 + (u64)pool.idx[slot]);
      for (;;) {
        StepPlan plan;
        const StepClass cls = classify_step<kLog, kBoxes>(sv, L, sp, ph, rng, st, plan);
        bool alive = false;
        TallyReq tr;
        if (cls == kVolume) alive = volume_step<kLog>(sv, L, sp, ph, rng, st, plan, tr);
        else if (cls == kSurface) alive = surface_step<kLog, kBoxes && kBoxSurface>(sv, L, sp, ph, rng, st, plan, tr);
        else if (cls == kExit) exit_step<kLog, kBoxes && kBoxSurface>(sv, L, sp, ph, st, plan, tr);
        else if (cls == kKill) kill_step<kLog>(sv, L, sp, ph, st, plan, tr);
        if (tr.sel >= 0) tally(sv, sink, ph, tr);
        if (!alive) break;
      }
      if (kLog && ph.log_ray >= 0) L.counts[ph.log_ray] = ph.nlog;
      pool.count[slot] = -1;
    }
  }
#endif
  if (S > 0 && tid < 32) {  // the last batch was published at the top of the final iteration
    if (tid == 0) {
      while (peek(pool.counters + kSvcAck) < (uint32_t)kSvcWarps) { }
      pool.counters[kSvcExit] = 1u;
      __threadfence_block();
    }
    __syncwarp();
    asm volatile("bar.arrive 2, %0;" ::"n"(kSvcWarps * 32 + 32) : "memory");
  }
#ifdef PVT_PROFILE_STAGES
#if PVT_PROFILE_STAGES == 2  // when CTAs finish: stats[4..7] = earliest start, earliest end, latest end, sum of ends (ns)
  if (tid == 0) {
    u64 now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    atomicMax(a.g_stats + 4, ~prof_start);  // minima as maxima of the complement (the slots start at zero)
    atomicMax(a.g_stats + 5, ~now);
    atomicMax(a.g_stats + 6, now);
    atomicAdd(a.g_stats + 7, now - prof_start);
  }
#else
  if (tid == 0)
    for (int k = 0; k < 4; ++k) atomicAdd(a.g_stats + 4 + k, (u64)prof[k]);
#endif
#endif
  }  // tracing warps
  LaneStats mine;  // a thread's own three words: written by nobody else
  if (tid < T) { mine.steps = pool.stats[tid]; mine.events = pool.stats[T + tid]; mine.rays = pool.stats[2 * T + tid]; }
  retire_cta(a, R, mine);
}

// =========================================================================================================
// trace_kernel: one photon per lane in registers

// shared memory layout of trace_kernel: [mbarrier 16 B][blob]
__host__ __device__ inline size_t trace_smem_bytes(int blob_words_in_smem) { return 16 + (size_t)blob_words_in_smem * 8; }

template <class Rng, int SW>
__global__ void __launch_bounds__(kTraceThreads) trace_kernel(const __grid_constant__ TraceArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sblob = reinterpret_cast<double*>(smem_raw + 16);
  if (a.scene_in_smem) stage_blob(sblob, a.blob, (uint32_t)a.blob_words * 8u, bar);
  const SceneView sv{a.scene_in_smem ? sblob : a.blob, &a.hdr};
  const int R = sv.hdr().n_recorders;
  const TallySink sink = cta_sink(a, R);
  const StepParams sp = a.sp;
  const LogColumns& L = a.log;

  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  PhotonT<SW> ph;
  Rng rng;
  ph.log_base = -1; ph.log_ray = -1; ph.nlog = 0;
  bool alive = false, exhausted = false;
  long long res_next = 0, res_end = 0;  // warp-uniform reservoir [res_next, res_end) of photon indices
  LaneStats st;

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
          if (a.has_rays()) {
            ph.p = (a.const_mask & 1u) ? V3{a.cpos[0], a.cpos[1], a.cpos[2]} : V3{a.pos[3 * idx], a.pos[3 * idx + 1], a.pos[3 * idx + 2]};
            ph.d = (a.const_mask & 2u) ? V3{a.cdir[0], a.cdir[1], a.cdir[2]} : V3{a.dir[3 * idx], a.dir[3 * idx + 1], a.dir[3 * idx + 2]};
            ph.wl = (a.const_mask & 4u) ? a.cwl : a.wl[idx];
          } else {
            emit_ray(sv, a.keys, a.first_index + idx, ph.p, ph.d, ph.wl);
          }
          rng.init(a.keys, (u64)a.first_index + (u64)idx);
          ph.log_ray = a.record_every > 0 ? sampled_ordinal(idx, a.record_every) : -1;
          ph.log_base = ph.log_ray < 0 ? -1 : (long long)ph.log_ray * sp.max_events;
          begin_photon<true>(ph, L, sp, st);
          alive = true;
        } else {
          exhausted = true;  // the global counter is past n: nothing will ever arrive
        }
      }
    }
    if (!__any_sync(kFullMask, alive)) break;
    if (alive) {
      StepPlan plan;
      const StepClass cls = classify_step<true>(sv, L, sp, ph, rng, st, plan);
      TallyReq tr;
      if (cls == kVolume) alive = volume_step<true>(sv, L, sp, ph, rng, st, plan, tr);
      else if (cls == kSurface) alive = surface_step<true>(sv, L, sp, ph, rng, st, plan, tr);
      else {
        if (cls == kExit) exit_step<true>(sv, L, sp, ph, st, plan, tr);
        else if (cls == kKill) kill_step<true>(sv, L, sp, ph, st, plan, tr);
        alive = false;
      }
      if (tr.sel >= 0) tally(sv, sink, ph, tr);
      if (!alive && ph.log_ray >= 0) {  // a sampled ray publishes its event count when it retires
        L.counts[ph.log_ray] = ph.nlog;
        ph.log_ray = -1; ph.log_base = -1;
      }
    }
  }
  retire_cta(a, R, st);
}

// ---- the intersect stage on its own ----------------------------------------------------------------------
// next_hit + find_container over a ray array (photon_tracer.py:26-109 == _kernel.pyx:666-714).  Algorithmic traffic
// per ray: position + direction in (48 B), t0 (8 B) and the three node ids packed into one word (4 B) out
// = 60 B; the unpacked form of the C ABI writes three int32 arrays instead (68 B).
//
// intersect_ring_kernel: persistent CTAs walk tiles of THREADS rays.  A tile's 2 x 24 THREADS bytes are fetched by
// TWO bulk async copies (cp.async.bulk -> SASS UBLKCP, the TMA engine's 1-D path) into a ring of STAGES tiles in
// shared memory, each stage with its own mbarrier: the copies of the next STAGES - 1 tiles are in flight while one
// is intersected, no register holds data in flight, and a thread reads ITS ray's six words from shared memory
// (stride 3 doubles: conflict free) -- the coalescing is the copy engine's business.  Only the node records of the
// blob are staged (header + nodes: the stage touches nothing else).  Streaming stores: every byte is written once.
__host__ __device__ constexpr int intersect_node_words(int n_nodes) { return kHeaderWords + n_nodes * kNodeWords; }
__host__ __device__ constexpr size_t intersect_ring_smem(int n_nodes, int threads, int stages) {
  return 16 + 8 * (size_t)stages + 8 /*pad to 16*/ + (size_t)intersect_node_words(n_nodes) * 8 + (size_t)stages * threads * 48;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar_a, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar_a), "r"(parity)
        : "memory");
  }
}

// (a node id that is "none" is -1, whose low byte is the 0xff of the packed form; with no hit at all, hit and container
// are -1 and t0 is the +inf the reduction started from)
__device__ __forceinline__ uint32_t pack_ids(const Nearest& nh) {
  const uint32_t lo = __byte_perm((uint32_t)nh.hit, (uint32_t)nh.container, 0x0040);  // bytes: hit, container, 0, 0 (of hit: byte 0 again)
  return (lo & 0xffffu) | (((uint32_t)nh.adjacent & 0xffu) << 16);
}

#ifndef PVT_INTERSECT_LEAN
#define PVT_INTERSECT_LEAN 1
#endif
constexpr bool kIntersectLean = PVT_INTERSECT_LEAN != 0;  // see nearest_surface_boxes<kLean>

template <bool kPacked>
__device__ __forceinline__ void store_hit(const Nearest& nh, long long i, double* t0, uint32_t* packed, int32_t* hit,
                                          int32_t* container, int32_t* adjacent) {
  __stcs(t0 + i, nh.t0);
  if (kPacked) {
    __stcs(packed + i, pack_ids(nh));
  } else {
    __stcs(hit + i, nh.hit);
    __stcs(container + i, nh.container);
    __stcs(adjacent + i, nh.adjacent);
  }
}

template <int THREADS, int STAGES, int kMinCtas, bool kBoxes, bool kPacked>
__global__ void __launch_bounds__(THREADS, kMinCtas)
    intersect_ring_kernel(const __grid_constant__ Header hdr, const double* blob, const double* pos, const double* dir,
                          long long n, double* t0, uint32_t* packed, int32_t* hit, int32_t* container, int32_t* adjacent) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar0 = reinterpret_cast<uint64_t*>(smem_raw);           // scene staging
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + 16);      // [STAGES]
  const int node_words = intersect_node_words(hdr.n_nodes);
  double* sblob = reinterpret_cast<double*>(smem_raw + 16 + 8 * STAGES + 8 * (STAGES & 1));
  double* ring = sblob + node_words;  // [STAGES][pos 3 THREADS | dir 3 THREADS]
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(full + s)));
  }
  stage_blob(sblob, blob, (uint32_t)node_words * 8u, bar0);  // (its barrier also publishes the inits above)
  const SceneView sv{sblob, &hdr};

  const long long tiles = n / THREADS;  // whole tiles go through the ring, the rest through plain loads below
  constexpr uint32_t kTileBytes = THREADS * 24u;
  auto issue = [&](long long tile, int s) {  // thread 0: both copies of `tile` into stage s
    const uint32_t bar_a = smem_addr(full + s), dst = smem_addr(ring + (size_t)s * THREADS * 6);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(2u * kTileBytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(pos + tile * THREADS * 3), "r"(kTileBytes), "r"(bar_a) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + kTileBytes),
                 "l"(dir + tile * THREADS * 3), "r"(kTileBytes), "r"(bar_a) : "memory");
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      const long long tile = (long long)blockIdx.x + (long long)s * gridDim.x;
      if (tile < tiles) issue(tile, s);
    }
  }
  int s = 0;
  uint32_t parity = 0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    mbar_wait(smem_addr(full + s), parity);
    const double* r = ring + (size_t)s * THREADS * 6;
    const V3 p = V3{r[3 * tid], r[3 * tid + 1], r[3 * tid + 2]};
    const V3 d = V3{r[THREADS * 3 + 3 * tid], r[THREADS * 3 + 3 * tid + 1], r[THREADS * 3 + 3 * tid + 2]};
    __syncthreads();  // every thread has its ray: the stage may be refilled
    if (tid == 0) {
      const long long next = tile + (long long)STAGES * gridDim.x;
      if (next < tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads before the async-proxy write
        issue(next, s);
      }
    }
    const Nearest nh = kBoxes ? nearest_surface_boxes<kIntersectLean>(sv, p, d) : nearest_surface(sv, p, d);
    store_hit<kPacked>(nh, tile * THREADS + tid, t0, packed, hit, container, adjacent);
    if (++s == STAGES) { s = 0; parity ^= 1u; }
  }
  // the last n % THREADS rays (a bulk copy moves multiples of 16 bytes from 16-byte aligned addresses only)
  if (blockIdx.x == (unsigned)(tiles % gridDim.x)) {
    const long long i = tiles * THREADS + tid;
    if (i < n) {
      const V3 p = V3{__ldcs(pos + 3 * i), __ldcs(pos + 3 * i + 1), __ldcs(pos + 3 * i + 2)};
      const V3 d = V3{__ldcs(dir + 3 * i), __ldcs(dir + 3 * i + 1), __ldcs(dir + 3 * i + 2)};
      const Nearest nh = kBoxes ? nearest_surface_boxes<kIntersectLean>(sv, p, d) : nearest_surface(sv, p, d);
      store_hit<kPacked>(nh, i, t0, packed, hit, container, adjacent);
    }
  }
}

// Plain-load form for ray arrays that are not 16-byte aligned (a bulk copy cannot start there) and for scenes whose
// node records do not fit beside the ring: one ray per thread, grid-stride.
template <bool kBoxes, bool kPacked>
__global__ void __launch_bounds__(256) intersect_plain_kernel(const __grid_constant__ Header hdr, const double* blob,
                                                              const double* pos, const double* dir, long long n, double* t0,
                                                              uint32_t* packed, int32_t* hit, int32_t* container,
                                                              int32_t* adjacent) {
  const SceneView sv{blob, &hdr};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const V3 p = V3{__ldcs(pos + 3 * i), __ldcs(pos + 3 * i + 1), __ldcs(pos + 3 * i + 2)};
    const V3 d = V3{__ldcs(dir + 3 * i), __ldcs(dir + 3 * i + 1), __ldcs(dir + 3 * i + 2)};
    const Nearest nh = kBoxes ? nearest_surface_boxes<kIntersectLean>(sv, p, d) : nearest_surface(sv, p, d);
    store_hit<kPacked>(nh, i, t0, packed, hit, container, adjacent);
  }
}

}  // namespace pvt
