// pvt_aux_kernels.cuh -- the small non-template kernels of libpvtrace_b200.so (emission into arrays, tally packing, the
// known-answer test launches).  Included by pvt_api.cu only.
#pragma once
#include "pvt_kernels.cuh"

namespace pvt {

__global__ void __launch_bounds__(256) emit_kernel(const double* blob, double* pos, double* dir, double* wl, long long n,
                                                   long long first_index, const __grid_constant__ RunSeed run) {
  const SceneView sv{blob};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    V3 p, d;
    double w;
    emit_ray(sv, run, first_index + i, p, d, w);
    pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
    dir[3 * i] = d.x; dir[3 * i + 1] = d.y; dir[3 * i + 2] = d.z;
    wl[i] = w;
  }
}

// fp64 FMA throughput of the chip, measured: eight independent chains per thread, every scheduler saturated.  The
// denominator of the tracer's COMPUTE roofline ("measure a DFMA microbenchmark before quoting a compute
// fraction"); 2 flops per FMA.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iterations) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-3 * (threadIdx.x + 32 * k);
  const double b = 0.999999, c = 1e-7;
  for (int i = 0; i < iterations; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
  }
  double sum = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) sum += a[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

// tallies <-> packed doubles (the buffer a multi-GPU caller all-reduces)
__global__ void pack_tallies_kernel(const u64* ints_a, int n_a, const double* sums, int n_s, const u64* bins, int n_b,
                                    double* packed) {
  const int total = n_a + n_s + n_b;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    double v;
    if (k < n_a) v = (double)ints_a[k];
    else if (k < n_a + n_s) v = sums[k - n_a];
    else v = (double)bins[k - n_a - n_s];
    packed[k] = v;
  }
}
__global__ void unpack_tallies_kernel(u64* ints_a, int n_a, double* sums, int n_s, u64* bins, int n_b,
                                      const double* packed) {
  const int total = n_a + n_s + n_b;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    const double v = packed[k];
    if (k < n_a) ints_a[k] = (u64)llrint(v);
    else if (k < n_a + n_s) sums[k - n_a] = v;
    else bins[k - n_a - n_s] = (u64)llrint(v);
  }
}

// ---- known-answer test kernels ----------------------------------------------------------------------------

__global__ void test_fresnel_kernel(long long n, const double* angle, const double* n1, const double* n2, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fresnel_R(angle[i], n1[i], n2[i]);
}
__global__ void test_reflect_kernel(long long n, const double* d, const double* nrm, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 r = mirror(V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]}, V3{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]});
  out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}
__global__ void test_refract_kernel(long long n, const double* d, const double* nrm, const double* n1, const double* n2,
                                    double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 dd = V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]};
  V3 nf = V3{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
  if (dot(nf, dd) < 0.0) nf = neg(nf);
  const V3 r = snell(dd, nf, n1[i], n2[i]);
  out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}
__global__ void test_intersect_kernel(long long n, const int32_t* gtype, const double* params, const double* o,
                                      const double* d, int32_t* nhit, double* ts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  const int k = roots(gtype[i], params + 4 * i, V3{o[3 * i], o[3 * i + 1], o[3 * i + 2]},
                      V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]}, t);
  nhit[i] = k;
  for (int j = 0; j < 4; ++j) ts[4 * i + j] = j < k ? t[j] : 0.0;
}
__global__ void test_normal_kernel(long long n, const int32_t* gtype, const double* params, const double* p, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 r = outward_normal(gtype[i], params + 4 * i, V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]});
  out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}
__global__ void test_interp_kernel(long long n, const double* x, int m, const double* xs, const double* ys, double inv_dx,
                                   double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = interp_hinted(x[i], xs, ys, m, inv_dx);
}
// op 0: log_lean(a), 1: div_newton(a, b), 2: rcp_newton(a), 3: sqrt_lean(a) --
// the lean forms of pvt_math.cuh against libm / IEEE arithmetic
__global__ void test_math_kernel(long long n, int op, const double* a, const double* b, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (op <= 2) { out[i] = op == 0 ? log_lean(a[i]) : (op == 1 ? div_newton(a[i], b[i]) : rcp_newton(a[i])); return; }
  out[i] = sqrt_lean(a[i]);
}
template <class Rng>
__global__ void test_rng_kernel(long long n_rays, int n_draws, u64 seed, long long first_index, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  Rng rng;
  const RunSeed run = make_run_seed(seed);
  rng.init(run, (u64)first_index + (u64)i);
  for (int k = 0; k < n_draws; ++k) out[i * n_draws + k] = rng.next();
}
template <class Rng>
__global__ void test_phase_kernel(long long n, int ptype, double prm, u64 seed, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rng rng;
  const RunSeed run = make_run_seed(seed);
  rng.init(run, (u64)i);
  const double g1 = rng.next(), g2 = rng.next();  // uniforms 0 and 1 of the ray's stream
  const V3 r = phase_direction(ptype, prm, g1, g2);
  out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}

}  // namespace pvt
