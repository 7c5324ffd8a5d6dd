// pvt_rng.cuh -- per-photon random streams.
//
//   Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), counter
//   based: fixed key, counter = (id lo, id hi, block, stream) with id = hash(seed) + first_index + i (see RunSeed):
//   runs with seeds s and s + 1 draw from far-apart windows of the counter space and share nothing, while bundles
//   of one run that continue each other's index range concatenate exactly (the reference's simulate_stream
//   contract, api.py:252-262).  One block gives two 53-bit uniforms.  Nothing is stored per photon but its index
//   and step number (see "Draw addressing" below).
//
//   xoshiro256+ seeded by splitmix64(seed + index): the reference's generator and per-ray seeding
//   (_kernel.pyx:75-113,1090), kept as a compatibility mode so per-ray histories can be compared with the
//   compiled reference kernel.
#pragma once
#include <stdint.h>

namespace pvt {

constexpr uint32_t kStreamTrace = 0u;
constexpr uint32_t kStreamEmit = 1u;

struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// The run's seed as the streams see it.  Philox keeps ONE compile-time key (its round keys are immediates of the
// instruction stream: a run-time key, even expanded on the host and read from the constant bank, measured 4 % of the
// trace kernel) and a run owns a WINDOW of the 64-bit counter space instead: photon i of the run draws from counter
// philox_base + i, philox_base = a 64-bit hash of the seed (splitmix64 finaliser).  Philox is a bijection of the
// counter, so two runs share a draw only if their windows overlap -- probability 2 n / 2^64 for runs of n photons --
// and seeds s and s + 1 land 2^63-ish apart.  Bundles of one run that continue each other's index range concatenate
// exactly.  The reference's xoshiro stream is seeded with seed + i as the reference does (_kernel.pyx:1090).
struct RunSeed {
  uint64_t seed;         // as given
  uint64_t philox_base;  // counter of photon 0 of the run
};
constexpr uint32_t kPhiloxKey0 = 0x50565442u;  // "PVTB"
constexpr uint32_t kPhiloxKey1 = 0x32303042u;  // "200B"
__host__ __device__ __forceinline__ uint64_t hash_seed(uint64_t seed) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ RunSeed make_run_seed(uint64_t seed) { return RunSeed{seed, hash_seed(seed)}; }

__host__ __device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

// uniform number k of stream `stream` of the photon with counter `id` (= philox_base + index): block k/2, half k%2
__host__ __device__ __forceinline__ double philox_uniform_at(uint64_t id, uint32_t stream, uint32_t k) {
  const U4 r = philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), k >> 1, stream}, kPhiloxKey0, kPhiloxKey1);
  return (k & 1u) ? u53(r.z, r.w) : u53(r.x, r.y);
}

// Draw addressing of the tracer.  Every random decision of a photon step has a fixed address
//   (photon index, step number, block, half)  ->  Philox counter (index_lo, index_hi, step * 8 + block, kStreamTrace)
// so a draw never depends on how many draws came before it: stages of the wavefront kernel can fetch their
// uniforms independently, and a skipped draw (e.g. no surface draw when the reflectivity is exactly zero,
// pvtrace/material/surface.py:231-240) does not shift the stream of the rays that follow.
enum {
  kBlockPath = 0,     // half 0: Beer-Lambert free path            half 1: surface reflect / transmit test
  kBlockAbsorb = 1,   // half 0: which component absorbs           half 1: quantum-yield test
  kBlockPhase = 2,    // both halves: phase-function direction
  kBlockEmit = 3,     // half 0: emission wavelength (gamma)       half 1: radiative / non-radiative delay
  kBlockLambert = 4,  // both halves: Lambertian reflection
  kBlocksPerStep = 8
};

struct PhiloxStream {
  static constexpr bool kAddressed = true;  // draws are fetched by address, in any order
  uint64_t id;    // the photon's counter: philox_base + photon index within the run (first_index + i)
  uint32_t step;  // 1-based trace-loop iteration of the photon (the reference's `count`)
  uint32_t k;     // sequential cursor, used only by next() (known-answer exports)
  __device__ __forceinline__ void init(const RunSeed& run, uint64_t index) { id = run.philox_base + index; step = 0; k = 0; }
  __device__ __forceinline__ void begin_step(uint32_t count) { step = count; }
  __device__ __forceinline__ U4 block(uint32_t b) const {
    return philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), step * kBlocksPerStep + b, kStreamTrace}, kPhiloxKey0, kPhiloxKey1);
  }
  __device__ __forceinline__ double one(uint32_t b, uint32_t half) {
    const U4 r = block(b);
    return half ? u53(r.z, r.w) : u53(r.x, r.y);
  }
  __device__ __forceinline__ void pair(uint32_t b, double& u0, double& u1) {
    const U4 r = block(b);
    u0 = u53(r.x, r.y);
    u1 = u53(r.z, r.w);
  }
  __device__ __forceinline__ double next() {
    return philox_uniform_at(id, kStreamTrace, k++);
  }
  // out-of-line copies for draws that almost never happen (delays, Lambertian mirrors): keeps the ten rounds
  // out of the hot instruction stream
  __device__ __noinline__ double one_rare(uint32_t b, uint32_t half) { return one(b, half); }
  __device__ __noinline__ void pair_rare(uint32_t b, double& u0, double& u1) { pair(b, u0, u1); }
};

// The reference's generator: draws are consumed in program order, addresses are ignored.
struct XoshiroStream {
  static constexpr bool kAddressed = false;  // draws must be consumed in the reference's program order
  uint64_t s0, s1, s2, s3;
  __device__ __forceinline__ static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  __device__ __forceinline__ void init(const RunSeed& run, uint64_t index) {
    uint64_t x = run.seed + index;  // the reference's per-ray seed, _kernel.pyx:1090
    s0 = splitmix(x); s1 = splitmix(x); s2 = splitmix(x); s3 = splitmix(x);
  }
  __device__ __forceinline__ void begin_step(uint32_t) {}
  __device__ __forceinline__ double next() {
    const uint64_t result = s0 + s3;
    const uint64_t t = s1 << 17;
    s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t;
    s3 = (s3 << 45) | (s3 >> 19);
    return (double)(result >> 11) * (1.0 / 9007199254740992.0);
  }
  __device__ __forceinline__ double one(uint32_t, uint32_t) { return next(); }
  __device__ __forceinline__ void pair(uint32_t, double& u0, double& u1) { u0 = next(); u1 = next(); }
  __device__ __forceinline__ double one_rare(uint32_t, uint32_t) { return next(); }
  __device__ __forceinline__ void pair_rare(uint32_t, double& u0, double& u1) { u0 = next(); u1 = next(); }
};

}  // namespace pvt
