// pvt_wave2.cu -- instantiations and launches of warp_wavefront_kernel (autonomous warps).
#include "pvt_common.cuh"
#include "pvt_launch.h"
#include "pvt_wave2.cuh"

namespace pvt {

static const Wave2Variant kWave2Table[] = {{16, 64}, {20, 64}, {16, 96}, {12, 64}, {8, 64}};
int wave2_variant_count() { return (int)(sizeof(kWave2Table) / sizeof(kWave2Table[0])); }
Wave2Variant wave2_variant(int k) { return kWave2Table[k]; }
size_t wave2_smem(const Wave2Variant& v, int blob_words, bool log) { return wave2_smem_bytes(blob_words, v.warps, v.slots, log); }

template <int W, int N>
static int wave2_case(const TraceArgs* args, bool boxes, bool log, int grid, size_t smem, cudaStream_t st) {
  if (!args) {  // configure: `smem` is the size of the logging instantiations (the larger pool)
    PVT_CUDA(cudaFuncSetAttribute(warp_wavefront_kernel<W, N, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PVT_CUDA(cudaFuncSetAttribute(warp_wavefront_kernel<W, N, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PVT_CUDA(cudaFuncSetAttribute(warp_wavefront_kernel<W, N, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PVT_CUDA(cudaFuncSetAttribute(warp_wavefront_kernel<W, N, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
  }
  if (boxes) {
    if (log) warp_wavefront_kernel<W, N, true, true><<<grid, W * 32, smem, st>>>(*args);
    else warp_wavefront_kernel<W, N, false, true><<<grid, W * 32, smem, st>>>(*args);
  } else {
    if (log) warp_wavefront_kernel<W, N, true, false><<<grid, W * 32, smem, st>>>(*args);
    else warp_wavefront_kernel<W, N, false, false><<<grid, W * 32, smem, st>>>(*args);
  }
  return 0;
}

#define PVT_WAVE2_CASE(W, N) if (v.warps == W && v.slots == N) return wave2_case<W, N>(args, boxes, log, grid, smem, st);
static int wave2_dispatch(const Wave2Variant& v, const TraceArgs* args, bool boxes, bool log, int grid, size_t smem,
                          cudaStream_t st) {
  PVT_WAVE2_CASE(16, 64) PVT_WAVE2_CASE(20, 64) PVT_WAVE2_CASE(16, 96) PVT_WAVE2_CASE(12, 64) PVT_WAVE2_CASE(8, 64)
  return fail("no warp-wavefront kernel variant for %d warps x %d slots", v.warps, v.slots);
}

int wave2_setup(const Wave2Variant& v, size_t smem_log) { return wave2_dispatch(v, nullptr, false, false, 0, smem_log, 0); }
int wave2_launch(const Wave2Variant& v, bool boxes, bool log, const TraceArgs& a, int grid, size_t smem, cudaStream_t st) {
  return wave2_dispatch(v, &a, boxes, log, grid, smem, st);
}

}  // namespace pvt
