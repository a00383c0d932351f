// Shared host/device helpers for libb2c.  See include/b2c.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "b2c.h"

namespace b2c {

// ---- error reporting (thread-local string behind b2c_last_error_string) -------------------
void set_error(const char *fmt, ...);

#define B2C_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::b2c::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

#define B2C_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      ::b2c::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                       __LINE__);                                                       \
      return B2C_E_CUDA;                                                                \
    }                                                                                   \
  } while (0)

int check_geom(const b2c_geom *g, bool allow_pitch = false);

// Opt-in to > 48 KB of dynamic shared memory ONCE per (kernel instantiation, device) instead of on every launch.
// `Kern` is the kernel itself (a non-type template argument), so every instantiation keeps its own per-device record.
constexpr int MAX_DEVICES = 64;
template <auto Kern>
inline cudaError_t set_max_smem(size_t bytes) {
  static std::atomic<int> granted[MAX_DEVICES];     // zero-initialised; bytes already granted on each device
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool tracked = dev >= 0 && dev < MAX_DEVICES;
  if (tracked && granted[dev].load(std::memory_order_acquire) >= (int)bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && tracked) granted[dev].store((int)bytes, std::memory_order_release);
  return e;
}

// SM count of the current device (cached per device ordinal).
inline int sm_count_current(int *out) {
  static std::atomic<int> cached[MAX_DEVICES];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  const bool tracked = dev >= 0 && dev < MAX_DEVICES;
  int n = tracked ? cached[dev].load(std::memory_order_acquire) : 0;
  if (!n) {
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    if (tracked) cached[dev].store(n, std::memory_order_release);
  }
  *out = n;
  return 0;
}

// ---- complex helpers (float2 = (re, im)) ---------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 cscale(float s, float2 a) { return make_float2(s * a.x, s * a.y); }
__device__ __forceinline__ float cabs2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }

// exp(j 2 pi t) for t in turns.  Range-reduce exactly (t - rint(t) is exact in fp32 for the
// |t| < 2 seen here), then the SFU sin/cos: abs error ~4e-7 on [-pi, pi].
__device__ __forceinline__ float2 cis_turns(float t) {
  float v = t - rintf(t);
  float s, c;
  __sincosf(6.283185307179586f * v, &s, &c);
  return make_float2(c, s);
}

// LS division y / (x + 1e-12): the epsilon lands on the real part of the complex pilot
// (src/baseline_estimators.py:40,110,171).
__device__ __forceinline__ float2 ls_divide(float2 y, float2 x) {
  float xr = x.x + 1e-12f, xi = x.y;
  float inv = 1.0f / fmaf(xr, xr, xi * xi);
  return make_float2(fmaf(y.x, xr, y.y * xi) * inv, fmaf(y.y, xr, -y.x * xi) * inv);
}

// ---- interpolation plan entry (16 B, see b2c_patterns in b2c.h) ----------------------------
// value = w0*h[i0] + w1*h[i1] + (1-w0-w1)*h[i2].  Resource elements outside the pilots' convex hull
// point all three indices at the zero slot h[np_max] that the kernels append to the pilot vector,
// which yields griddata's fill_value = 0.0 exactly and needs no flag test.
struct PlanTap {
  int i0, i1, i2;
  float w0, w1, w2;
};
__device__ __forceinline__ PlanTap plan_decode(uint4 raw) {
  PlanTap p;
  p.i0 = raw.x & 0xffffu;
  p.i1 = raw.x >> 16;
  p.i2 = raw.y & 0xffffu;
  p.w0 = __uint_as_float(raw.z);
  p.w1 = __uint_as_float(raw.w);
  p.w2 = 1.0f - p.w0 - p.w1;
  return p;
}
__device__ __forceinline__ float2 plan_apply(const PlanTap &p, const float2 *__restrict__ hp) {
  float2 a = hp[p.i0], b = hp[p.i1], c = hp[p.i2];
  return make_float2(fmaf(p.w0, a.x, fmaf(p.w1, b.x, p.w2 * c.x)),
                     fmaf(p.w0, a.y, fmaf(p.w1, b.y, p.w2 * c.y)));
}

// ---- reductions ------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in every thread.  `scratch` holds >= 33 floats.
__device__ __forceinline__ float block_sum(float v, float *scratch) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // scratch may still be in use by a previous call
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nwarp ? scratch[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// streaming 8-byte store (outputs are written once and not re-read by this kernel)
__device__ __forceinline__ void st_stream(float2 *p, float2 v) { __stcs(p, v); }

}  // namespace b2c
