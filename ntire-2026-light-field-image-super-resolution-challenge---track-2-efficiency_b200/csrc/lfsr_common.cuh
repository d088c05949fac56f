// Shared helpers for liblfsr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include "../../include/lfsr.h"

namespace lfsr {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return LFSR_ERR_CUDA;
  }
  return LFSR_OK;
}

// CUDA function attributes (and SM counts) belong to a DEVICE: a process that later runs on a second GPU must opt every
// kernel into > 48 KB of dynamic shared memory there too, so one-time setup is tracked per device and its return code is
// checked. Usage:  static DevOnce once;  if (once.need()) { <set attributes, return on error>;  once.done(); }
struct DevOnce {
  std::atomic<uint64_t> mask{0};
  int dev = 0;
  bool need() {
    int d = 0;
    cudaGetDevice(&d);
    return !(mask.load(std::memory_order_acquire) & (1ull << (d & 63)));
  }
  void done() {
    int d = 0;
    cudaGetDevice(&d);
    mask.fetch_or(1ull << (d & 63), std::memory_order_release);
  }
};
template <class K>
inline int opt_in_smem(K kernel, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %d): %s", what, bytes, cudaGetErrorString(e));
    return LFSR_ERR_CUDA;
  }
  return LFSR_OK;
}
// SM count of the CURRENT device (cached per device ordinal)
inline int sm_count_current() {
  static std::atomic<int> cache[64];
  int d = 0;
  cudaGetDevice(&d);
  int v = cache[d & 63].load(std::memory_order_relaxed);
  if (v <= 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d);
    cache[d & 63].store(v, std::memory_order_relaxed);
  }
  return v;
}
// Experiment switches (profiles/ probes) are read from the environment ONLY in the probe build (-DLFSR_DEBUG_HOOKS,
// liblfsr_probe.so): the product library ignores them, so a stray variable can neither steer nor corrupt a product run.
inline const char* dbg_env(const char* name) {
#ifdef LFSR_DEBUG_HOOKS
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

#define LFSR_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::lfsr::set_error(__VA_ARGS__);      \
      return LFSR_ERR_INVALID;             \
    }                                      \
  } while (0)

inline bool tensor_ok(const lfsr_tensor* t) {
  return t && t->ptr && t->n > 0 && t->h > 0 && t->w > 0 && t->c > 0 && t->ld >= t->c;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// device-side view of lfsr_tensor
struct TView {
  float* p;
  int n, h, w, c, ld;
  __device__ __forceinline__ size_t pix(int in_, int y, int x) const {
    return ((size_t)((size_t)in_ * h + y) * w + x) * (size_t)ld;
  }
};
inline TView view_of(const lfsr_tensor* t) {
  TView v;
  v.p = (float*)t->ptr; v.n = t->n; v.h = t->h; v.w = t->w; v.c = t->c; v.ld = t->ld;
  return v;
}
inline TView null_view() { TView v; v.p = nullptr; v.n = v.h = v.w = v.c = v.ld = 0; return v; }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case LFSR_ACT_RELU: return v > 0.f ? v : 0.f;
    case LFSR_ACT_LRELU: return v > 0.f ? v : v * slope;
    case LFSR_ACT_SIGMOID: return __fdividef(1.f, 1.f + __expf(-v));
    case LFSR_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    case LFSR_ACT_SILU: return __fdividef(v, 1.f + __expf(-v));
    default: return v;
  }
}

// packed 2 x fp32 FMA (FFMA2 on sm_100): each half is an ordinary fma.rn, so results equal two scalar FMAs bit for bit
// while the issue-bound CUDA-core kernels retire half as many instructions
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// MacPI (i*A+u, j*A+v) logical coordinate -> SAI (u*hh+i, v*ww+j) storage coordinate,
// H = A*hh, W = A*ww  (DistgSSR.py:134-155)
__device__ __forceinline__ void macpi_to_sai(int y, int x, int A, int H, int W, int& sy, int& sx) {
  int hh = H / A, ww = W / A;
  int i = y / A, u = y - i * A;
  int j = x / A, v = x - j * A;
  sy = u * hh + i;
  sx = v * ww + j;
}

}  // namespace lfsr
