// pvt_intersect.cu -- instantiations and launches of the intersect stage (intersect_ring_kernel, intersect_plain_kernel).
#include <stdint.h>
#include <stdlib.h>

#include "pvt_common.cuh"
#include "pvt_kernels.cuh"
#include "pvt_launch.h"

namespace pvt {

// (threads per CTA, ring stages, CTAs per SM); the first is the default, PVT_INTERSECT_VARIANT picks another by index
struct RingShape { int threads, stages, ctas; };
static const RingShape kRingShapes[] = {{256, 4, 3}, {256, 6, 3}, {128, 5, 6}, {256, 3, 4}, {128, 8, 4}, {512, 4, 1}};
constexpr int kRingShapeCount = (int)(sizeof(kRingShapes) / sizeof(kRingShapes[0]));

template <int THREADS, int STAGES, int CTAS>
static int ring_launch(bool boxes, bool packed, const Header& hdr, const double* blob, const double* pos, const double* dir,
                       long long n, double* t0, uint32_t* ids, int32_t* hit, int32_t* container, int32_t* adjacent,
                       int sm_count, cudaStream_t st) {
  const size_t smem = intersect_ring_smem(hdr.n_nodes, THREADS, STAGES);
  const long long tiles = (n + THREADS - 1) / THREADS;
  const long long resident = (long long)sm_count * CTAS;
  const int grid = (int)(tiles < resident ? (tiles < 1 ? 1 : tiles) : resident);
#define PVT_RING_GO(BX, PK)                                                                                             \
  do {                                                                                                                  \
    auto kernel = intersect_ring_kernel<THREADS, STAGES, CTAS, BX, PK>;                                                 \
    static bool configured = false;                                                                                     \
    if (!configured) {                                                                                                  \
      PVT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));                  \
      configured = true;                                                                                                \
    }                                                                                                                   \
    kernel<<<grid, THREADS, smem, st>>>(hdr, blob, pos, dir, n, t0, ids, hit, container, adjacent);                     \
  } while (0)
  if (boxes) { if (packed) PVT_RING_GO(true, true); else PVT_RING_GO(true, false); }
  else { if (packed) PVT_RING_GO(false, true); else PVT_RING_GO(false, false); }
#undef PVT_RING_GO
  PVT_CUDA(cudaGetLastError());
  return 0;
}

int intersect_launch(bool boxes, bool packed, const Header& hdr, const double* blob, const double* pos, const double* dir,
                     long long n, double* t0, uint32_t* ids, int32_t* hit, int32_t* container, int32_t* adjacent,
                     int sm_count, cudaStream_t st) {
  int which = 0;
  if (const char* env = getenv("PVT_INTERSECT_VARIANT")) which = atoi(env);
  const bool aligned = (((uintptr_t)pos | (uintptr_t)dir) & 15u) == 0;
  // (the "configured" flags above are per process, fine for one device type; the attribute is per function and device,
  // so a second device of a different kind would need it set again -- every B200 of a node is the same)
  if (which < 0 || which >= kRingShapeCount || !aligned ||
      intersect_ring_smem(hdr.n_nodes, kRingShapes[which].threads, kRingShapes[which].stages) * kRingShapes[which].ctas > 220 * 1024) {
    const long long blocks = (n + 255) / 256, cap = (long long)sm_count * 8;
    const int grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
    if (boxes) {
      if (packed) intersect_plain_kernel<true, true><<<grid, 256, 0, st>>>(hdr, blob, pos, dir, n, t0, ids, hit, container, adjacent);
      else intersect_plain_kernel<true, false><<<grid, 256, 0, st>>>(hdr, blob, pos, dir, n, t0, ids, hit, container, adjacent);
    } else {
      if (packed) intersect_plain_kernel<false, true><<<grid, 256, 0, st>>>(hdr, blob, pos, dir, n, t0, ids, hit, container, adjacent);
      else intersect_plain_kernel<false, false><<<grid, 256, 0, st>>>(hdr, blob, pos, dir, n, t0, ids, hit, container, adjacent);
    }
    PVT_CUDA(cudaGetLastError());
    return 0;
  }
#define PVT_RING_CASE(K, T, S, C) \
  if (which == K) return ring_launch<T, S, C>(boxes, packed, hdr, blob, pos, dir, n, t0, ids, hit, container, adjacent, sm_count, st);
  PVT_RING_CASE(0, 256, 4, 3) PVT_RING_CASE(1, 256, 6, 3) PVT_RING_CASE(2, 128, 5, 6) PVT_RING_CASE(3, 256, 3, 4)
  PVT_RING_CASE(4, 128, 8, 4) PVT_RING_CASE(5, 512, 4, 1)
#undef PVT_RING_CASE
  return fail("no intersect kernel variant %d", which);
}

}  // namespace pvt
