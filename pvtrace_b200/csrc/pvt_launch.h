// pvt_launch.h -- host-side launch entry points of the tracer kernels.  Each family of template instantiations lives in
// its own translation unit (pvt_wave.cu, pvt_reg.cu) so that they compile in parallel; pvt_api.cu calls
// them through these functions.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace pvt {

struct TraceArgs;

// wavefront_kernel<T, P, B, kLog, kBoxes, S>: shapes (threads, pool slots, CTAs per SM) of the variant table
struct WaveVariant { int threads, pool, ctas; };
int wave_variant_count();
WaveVariant wave_variant(int k);
// raises the dynamic shared-memory limit of every instantiation of the shape; non-zero (with pvt_last_error set) on failure
int wave_setup(const WaveVariant& v, size_t smem);
// service: 0 or wave_service_threads() (only the {512, 1024, 1} shape has service-warp instantiations)
int wave_service_threads();
int wave_launch(const WaveVariant& v, int service, bool boxes, bool log, const TraceArgs& a, int grid, size_t smem,
                cudaStream_t st);

// trace_kernel<Rng, SW>: which = (xoshiro ? 2 : 0) + (more than 64 recorders ? 1 : 0)
int reg_occupancy(int which, size_t smem, int* blocks_per_sm);
int reg_launch(int which, const TraceArgs& a, int grid, size_t smem, cudaStream_t st);

}  // namespace pvt
