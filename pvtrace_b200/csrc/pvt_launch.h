// pvt_launch.h -- host-side launch entry points of the tracer kernels.  Each family of template instantiations lives in
// its own translation unit (pvt_wave.cu, pvt_reg.cu) so that they compile in parallel; pvt_api.cu calls
// them through these functions.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

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

// the intersect stage on its own: ring of bulk copies when the arrays are 16-byte aligned, plain loads otherwise.
// packed: node ids as one word per ray (hit | container << 8 | adjacent << 16, 0xff = none) into `ids`; else three int32 arrays
struct Header;
int intersect_launch(bool boxes, bool packed, const Header& hdr, const double* blob, const double* pos, const double* dir,
                     long long n, double* t0, uint32_t* ids, int32_t* hit, int32_t* container, int32_t* adjacent,
                     int sm_count, cudaStream_t st);

// trace_kernel<Rng, SW>: which = (xoshiro ? 2 : 0) + (more than 64 recorders ? 1 : 0)
int reg_occupancy(int which, size_t smem, int* blocks_per_sm);
int reg_launch(int which, const TraceArgs& a, int grid, size_t smem, cudaStream_t st);

}  // namespace pvt
