// pvt_reg.cu -- instantiations and launches of trace_kernel (one photon per lane in registers).
#include "pvt_common.cuh"
#include "pvt_kernels.cuh"
#include "pvt_launch.h"

namespace pvt {

template <class K>
static int occupancy_of(K kernel, size_t smem, int* blocks) {
  if (smem > 48 * 1024)
    PVT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PVT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kernel, kTraceThreads, smem));
  if (*blocks < 1) return fail("trace kernel cannot be resident with %zu bytes of shared memory", smem);
  return 0;
}

int reg_occupancy(int which, size_t smem, int* blocks) {
  switch (which) {
    case 0: return occupancy_of(trace_kernel<PhiloxStream, 2>, smem, blocks);
    case 1: return occupancy_of(trace_kernel<PhiloxStream, 8>, smem, blocks);
    case 2: return occupancy_of(trace_kernel<XoshiroStream, 2>, smem, blocks);
    default: return occupancy_of(trace_kernel<XoshiroStream, 8>, smem, blocks);
  }
}

int reg_launch(int which, const TraceArgs& a, int grid, size_t smem, cudaStream_t st) {
  switch (which) {
    case 0: trace_kernel<PhiloxStream, 2><<<grid, kTraceThreads, smem, st>>>(a); break;
    case 1: trace_kernel<PhiloxStream, 8><<<grid, kTraceThreads, smem, st>>>(a); break;
    case 2: trace_kernel<XoshiroStream, 2><<<grid, kTraceThreads, smem, st>>>(a); break;
    default: trace_kernel<XoshiroStream, 8><<<grid, kTraceThreads, smem, st>>>(a); break;
  }
  return 0;
}

}  // namespace pvt
