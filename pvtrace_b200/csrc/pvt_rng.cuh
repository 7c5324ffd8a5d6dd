// pvt_rng.cuh -- per-photon random streams.
//
//   Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), counter
//   based: counter = (ray id lo, ray id hi, block, stream), fixed key.  One block gives two 53-bit uniforms, so
//   draw k of a ray lives in block k/2 -- random access, nothing to store but the draw index.  The ray id is
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

// draw number k of stream `stream` of ray `id`
__host__ __device__ __forceinline__ double philox_uniform_at(uint64_t id, uint32_t stream, uint32_t k) {
  const U4 r = philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), k >> 1, stream}, kPhiloxKey0, kPhiloxKey1);
  return (k & 1u) ? u53(r.z, r.w) : u53(r.x, r.y);
}

struct PhiloxStream {
  uint64_t id;
  uint32_t k;     // next draw index
  double spare;   // second uniform of the current block, valid when k is odd
  __device__ __forceinline__ void init(uint64_t ray_id) { id = ray_id; k = 0; spare = 0.0; }
  __device__ __forceinline__ double next() {
    if (k & 1u) { ++k; return spare; }
    const U4 r = philox4x32_10(U4{(uint32_t)id, (uint32_t)(id >> 32), k >> 1, kStreamTrace}, kPhiloxKey0, kPhiloxKey1);
    ++k;
    spare = u53(r.z, r.w);
    return u53(r.x, r.y);
  }
};

struct XoshiroStream {
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
  __device__ __forceinline__ double next() {
    const uint64_t result = s0 + s3;
    const uint64_t t = s1 << 17;
    s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t;
    s3 = (s3 << 45) | (s3 >> 19);
    return (double)(result >> 11) * (1.0 / 9007199254740992.0);
  }
};

}  // namespace pvt
