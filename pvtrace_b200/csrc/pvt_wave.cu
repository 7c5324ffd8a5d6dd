// pvt_wave.cu -- instantiations and launches of wavefront_kernel (two-stage CTA wavefront with service warps).
#include "pvt_common.cuh"
#include "pvt_kernels.cuh"
#include "pvt_launch.h"

#ifndef PVT_SVC_THREADS
#define PVT_SVC_THREADS 128  // service threads per CTA of the {512, 1024, 1} shape (64: two service warps, more registers per tracer)
#endif

namespace pvt {

int wave_service_threads() { return PVT_SVC_THREADS; }

// (threads, pool slots, resident CTAs per SM), preferred first; the first that fits the scene's shared memory wins
#ifdef PVT_SWEEP  // extra shapes for tools/sweep.sh experiments (selected with PVT_WAVEFRONT_* in the environment)
constexpr int kWaveVariants = 11;
#else
constexpr int kWaveVariants = 7;
#endif
static const WaveVariant kWaveTable[kWaveVariants] = {{512, 1024, 1}, {480, 960, 1}, {448, 896, 1}, {512, 768, 1},
                                                      {512, 512, 1},  {384, 768, 1}, {256, 512, 1},
#ifdef PVT_SWEEP
                                                      {256, 512, 2}, {640, 1280, 1}, {768, 768, 1}, {1024, 1024, 1},
#endif
};
int wave_variant_count() { return kWaveVariants; }
WaveVariant wave_variant(int k) { return kWaveTable[k]; }

template <class K>
static int wave_attr(K kernel, size_t smem) {
  PVT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

// launches (or, with args == nullptr, only configures) one shape
template <int T, int P, int B>
static int wave_case(const TraceArgs* args, int service, bool boxes, bool log, int grid, size_t smem, cudaStream_t st) {
  constexpr int TS = (T == 512 && P == 1024 && B == 1) ? PVT_SVC_THREADS : 0;
  if (!args) {
    PVT_TRY(wave_attr(wavefront_kernel<T, P, B, false, false>, smem));
    PVT_TRY(wave_attr(wavefront_kernel<T, P, B, false, true>, smem));
    PVT_TRY(wave_attr(wavefront_kernel<T, P, B, true, false>, smem));
    PVT_TRY(wave_attr(wavefront_kernel<T, P, B, true, true>, smem));
    if (TS > 0) {
      PVT_TRY(wave_attr(wavefront_kernel<T, P, B, false, false, TS>, smem));
      PVT_TRY(wave_attr(wavefront_kernel<T, P, B, false, true, TS>, smem));
      PVT_TRY(wave_attr(wavefront_kernel<T, P, B, true, false, TS>, smem));
      PVT_TRY(wave_attr(wavefront_kernel<T, P, B, true, true, TS>, smem));
    }
    return 0;
  }
  if (TS > 0 && service > 0) {
    if (boxes) {
      if (log) wavefront_kernel<T, P, B, true, true, TS><<<grid, T + TS, smem, st>>>(*args);
      else wavefront_kernel<T, P, B, false, true, TS><<<grid, T + TS, smem, st>>>(*args);
    } else {
      if (log) wavefront_kernel<T, P, B, true, false, TS><<<grid, T + TS, smem, st>>>(*args);
      else wavefront_kernel<T, P, B, false, false, TS><<<grid, T + TS, smem, st>>>(*args);
    }
    return 0;
  }
  if (boxes) {
    if (log) wavefront_kernel<T, P, B, true, true><<<grid, T, smem, st>>>(*args);
    else wavefront_kernel<T, P, B, false, true><<<grid, T, smem, st>>>(*args);
  } else {
    if (log) wavefront_kernel<T, P, B, true, false><<<grid, T, smem, st>>>(*args);
    else wavefront_kernel<T, P, B, false, false><<<grid, T, smem, st>>>(*args);
  }
  return 0;
}

#define PVT_WAVE_CASE(T, P, B) \
  if (v.threads == T && v.pool == P && v.ctas == B) return wave_case<T, P, B>(args, service, boxes, log, grid, smem, st);
static int wave_dispatch(const WaveVariant& v, const TraceArgs* args, int service, bool boxes, bool log, int grid,
                         size_t smem, cudaStream_t st) {
  PVT_WAVE_CASE(512, 1024, 1) PVT_WAVE_CASE(480, 960, 1) PVT_WAVE_CASE(448, 896, 1) PVT_WAVE_CASE(512, 768, 1)
  PVT_WAVE_CASE(512, 512, 1) PVT_WAVE_CASE(384, 768, 1) PVT_WAVE_CASE(256, 512, 1)
#ifdef PVT_SWEEP
  PVT_WAVE_CASE(256, 512, 2) PVT_WAVE_CASE(640, 1280, 1) PVT_WAVE_CASE(768, 768, 1) PVT_WAVE_CASE(1024, 1024, 1)
#endif
  return fail("no wavefront kernel variant for %d threads / %d slots x %d CTAs", v.threads, v.pool, v.ctas);
}

int wave_setup(const WaveVariant& v, size_t smem) { return wave_dispatch(v, nullptr, 0, false, false, 0, smem, 0); }
int wave_launch(const WaveVariant& v, int service, bool boxes, bool log, const TraceArgs& a, int grid, size_t smem,
                cudaStream_t st) {
  return wave_dispatch(v, &a, service, boxes, log, grid, smem, st);
}

}  // namespace pvt
