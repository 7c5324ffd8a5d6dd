// pvt_rng.cuh -- per-photon random streams.
//
//   Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), counter
//   based: counter = (ray id lo, ray id hi, block, stream), fixed key.  One block gives two 53-bit uniforms.
//   Nothing is stored per photon but its id and step number (see "Draw addressing" below).  The ray id is
//   seed + first_index + i, the reference's per-ray seed (pvtrace/engine/_kernel.pyx:1090), which keeps the
//   reference's "bundles with consecutive seed offsets concatenate exactly" contract (api.py:252-262).
//
//   xoshiro256+ seeded by splitmix64: the reference's generator (_kernel.pyx:75-113), kept as a compatibility
//   mode so per-ray histories can be compared with the compiled reference kernel.
#pragma once
#include <stdint.h>

namespace pvt {

constexpr uint32_t kPhiloxKey0 = 0x50565442u;  // "PVTB"
constexpr uint32_t kPhiloxKey1 = 0x32303042u;  // "200B"
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

__host__ __device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

// uniform number k of stream `stream` of ray `id` (random access): block k/2, half k%2
__host__ __device__ __forceinline__ double philox_uniform_at(uint64_t id, uint32_t stream, uint32_t k) {
  const U4 r = philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), k >> 1, stream}, kPhiloxKey0, kPhiloxKey1);
  return (k & 1u) ? u53(r.z, r.w) : u53(r.x, r.y);
}

// Draw addressing of the tracer.  Every random decision of a photon step has a fixed address
//   (photon id, step number, block, half)  ->  Philox counter (id_lo, id_hi, step * 8 + block, kStreamTrace)
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
  uint64_t id;
  uint32_t step;  // 1-based trace-loop iteration of the photon (the reference's `count`)
  uint32_t k;     // sequential cursor, used only by next() (known-answer exports)
  __device__ __forceinline__ void init(uint64_t ray_id) { id = ray_id; step = 0; k = 0; }
  __device__ __forceinline__ void begin_step(uint32_t count) { step = count; }
  __device__ __forceinline__ U4 block(uint32_t b) const {
    return philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), step * kBlocksPerStep + b, kStreamTrace}, kPhiloxKey0,
                         kPhiloxKey1);
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
  __device__ __forceinline__ double next() { return philox_uniform_at(id, kStreamTrace, k++); }
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
  __device__ __forceinline__ void init(uint64_t ray_id) {
    uint64_t x = ray_id;
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
