// pvt_common.cuh -- error plumbing shared by the translation units of libpvtrace_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

namespace pvt {

// thread-local message returned by pvt_last_error()
char* error_buffer();
int fail(const char* fmt, ...);

#define PVT_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t pvt_err__ = (expr);                                                                 \
    if (pvt_err__ != cudaSuccess)                                                                   \
      return ::pvt::fail("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(pvt_err__)); \
  } while (0)

#define PVT_TRY(expr)            \
  do {                           \
    int pvt_rc__ = (expr);       \
    if (pvt_rc__ != 0) return pvt_rc__; \
  } while (0)

template <class T>
struct DeviceBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }  // (owners set their device current first; at process exit the runtime may be gone: harmless)
  int reserve(size_t n) {
    if (n <= count && ptr) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
    if (n == 0) return 0;
    cudaError_t e = cudaMalloc((void**)&ptr, n * sizeof(T));
    if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) -> %s", n * sizeof(T), cudaGetErrorString(e));
    count = n;
    return 0;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
};

}  // namespace pvt
