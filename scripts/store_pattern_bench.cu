// Diagnostic micro-benchmark (not part of libb2c): replays the slot kernel's STORE PATTERN with no compute,
// to find what the memory system gives this pattern -- the ceiling the fused kernel can reach -- and how
// it moves with row alignment, store policy and CTA shape.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/build/store_pattern_bench \
//        scripts/store_pattern_bench.cu && scripts/build/store_pattern_bench [B]
//
// Pattern (b2c_slot.cu, FAST 4x4 path): one CTA per (slot, rx), 320 threads, thread t owns bins
// k0 = 299 - t and k1 = 300 + t; per symbol it stores H_true / H_ls / H_mmse rows for 4 tx, one rx row, and
// (rx == 0) 4 tx-grid rows; rows are `pitch` complex64 apart (599 in the product).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("%s: %s\n", #x, cudaGetErrorString(e));                               \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int NSYM = 14, NRX = 4, NTX = 4, NSC = 599, HALF = 300;

template <int POLICY>
__device__ __forceinline__ void st(float2 *p, float2 v) {
  if (POLICY == 0) __stcs(p, v);
  else if (POLICY == 1) *p = v;
  else if (POLICY == 2) __stwt(p, v);
  else __stcg(p, v);
}

// MODE 0: the product's mirror-bin ownership; MODE 1: thread t owns bins t and t + 320 (ascending)
template <int POLICY, int MODE>
__global__ void __launch_bounds__(320, 2) pattern_kernel(float2 *H, float2 *L, float2 *M, float2 *R, float2 *T, int pitch,
                                                         int spin) {
  const int b = blockIdx.x / NRX, rx = blockIdx.x % NRX, t = threadIdx.x;
  int k0, k1;
  bool v0, v1;
  if (MODE == 0) {
    v0 = t < HALF, v1 = t < HALF - 1;
    k0 = v0 ? HALF - 1 - t : 0, k1 = v1 ? HALF + t : 0;
  } else {
    v0 = true, v1 = t + 320 < NSC;
    k0 = t, k1 = v1 ? t + 320 : 0;
  }
  const int64_t slot_h = (int64_t)NSYM * NRX * NTX * pitch, slot_r = (int64_t)NSYM * NRX * pitch,
                slot_t = (int64_t)NSYM * NTX * pitch;
  float2 *pH = H + b * slot_h + rx * NTX * pitch, *pL = L + b * slot_h + rx * NTX * pitch,
         *pM = M + b * slot_h + rx * NTX * pitch, *pR = R + b * slot_r + rx * pitch, *pT = T + b * slot_t;
  float2 v = make_float2((float)t, (float)b);
  for (int s = 0; s < NSYM; ++s) {
    for (int i = 0; i < spin; ++i) v.x = fmaf(v.x, 1.0000001f, 1e-9f);      // stand-in for the compute between stores
#pragma unroll
    for (int tx = 0; tx < NTX; ++tx) {
      if (v0) st<POLICY>(pH + tx * pitch + k0, v);
      if (v1) st<POLICY>(pH + tx * pitch + k1, v);
      if (v0) st<POLICY>(pL + tx * pitch + k0, v);
      if (v1) st<POLICY>(pL + tx * pitch + k1, v);
      if (v0) st<POLICY>(pM + tx * pitch + k0, v);
      if (v1) st<POLICY>(pM + tx * pitch + k1, v);
    }
    if (v0) st<POLICY>(pR + k0, v);
    if (v1) st<POLICY>(pR + k1, v);
    if (rx == 0) {
#pragma unroll
      for (int tx = 0; tx < NTX; ++tx) {
        if (v0) st<POLICY>(pT + tx * pitch + k0, v);
        if (v1) st<POLICY>(pT + tx * pitch + k1, v);
      }
    }
    pH += NRX * NTX * pitch, pL += NRX * NTX * pitch, pM += NRX * NTX * pitch, pR += NRX * pitch, pT += NTX * pitch;
  }
}

// One CTA per (slot, symbol-pair ...) alternative: a CTA owns ALL rx of a slot for one symbol at a time, so that it
// writes 16 consecutive rows (76 KB contiguous) per array per symbol.
template <int POLICY>
__global__ void __launch_bounds__(320, 2) pattern_allrx_kernel(float2 *H, float2 *L, float2 *M, float2 *R, float2 *T, int pitch) {
  const int b = blockIdx.x, t = threadIdx.x;
  const bool v1 = t + 320 < NSC;
  const int k0 = t, k1 = v1 ? t + 320 : 0;
  const int64_t slot_h = (int64_t)NSYM * NRX * NTX * pitch, slot_r = (int64_t)NSYM * NRX * pitch,
                slot_t = (int64_t)NSYM * NTX * pitch;
  float2 *pH = H + b * slot_h, *pL = L + b * slot_h, *pM = M + b * slot_h, *pR = R + b * slot_r, *pT = T + b * slot_t;
  const float2 v = make_float2((float)t, (float)b);
  for (int s = 0; s < NSYM; ++s) {
#pragma unroll 4
    for (int r = 0; r < NRX * NTX; ++r) {
      st<POLICY>(pH + r * pitch + k0, v);
      if (v1) st<POLICY>(pH + r * pitch + k1, v);
      st<POLICY>(pL + r * pitch + k0, v);
      if (v1) st<POLICY>(pL + r * pitch + k1, v);
      st<POLICY>(pM + r * pitch + k0, v);
      if (v1) st<POLICY>(pM + r * pitch + k1, v);
    }
#pragma unroll
    for (int r = 0; r < NRX; ++r) {
      st<POLICY>(pR + r * pitch + k0, v);
      if (v1) st<POLICY>(pR + r * pitch + k1, v);
      st<POLICY>(pT + r * pitch + k0, v);
      if (v1) st<POLICY>(pT + r * pitch + k1, v);
    }
    pH += NRX * NTX * pitch, pL += NRX * NTX * pitch, pM += NRX * NTX * pitch, pR += NRX * pitch, pT += NTX * pitch;
  }
}

// fill variants: 8-byte stores; CTA-private contiguous regions (each CTA streams `chunk` bytes at a time from its own
// region, like one array of the product pattern but without the row structure)
__global__ void fill8_kernel(float2 *p, int64_t n) {
  const float2 v = make_float2(1.f, 2.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) __stcs(p + i, v);
}
template <typename VT>
__global__ void __launch_bounds__(320, 2) private_region_kernel(VT *p, int64_t per_cta) {   // per_cta in elements
  VT v;
  memset(&v, 0, sizeof(v));
  VT *q = p + (int64_t)blockIdx.x * per_cta;
  for (int64_t i = threadIdx.x; i < per_cta; i += blockDim.x) __stcs(q + i, v);
}

// The product pattern with 16-byte stores: thread t owns the adjacent bins (2t, 2t+1) of an ALIGNED row (pitch 600),
// 300 active threads, one STG.128 per row.
template <int POLICY>
__global__ void __launch_bounds__(320, 2) pattern128_kernel(float4 *H, float4 *L, float4 *M, float4 *R, float4 *T, int pitch2) {
  const int b = blockIdx.x / NRX, rx = blockIdx.x % NRX, t = threadIdx.x;
  if (t >= 300) return;
  const int64_t slot_h = (int64_t)NSYM * NRX * NTX * pitch2, slot_r = (int64_t)NSYM * NRX * pitch2,
                slot_t = (int64_t)NSYM * NTX * pitch2;
  float4 *pH = H + b * slot_h + rx * NTX * pitch2 + t, *pL = L + b * slot_h + rx * NTX * pitch2 + t,
         *pM = M + b * slot_h + rx * NTX * pitch2 + t, *pR = R + b * slot_r + rx * pitch2 + t, *pT = T + b * slot_t + t;
  const float4 v = make_float4((float)t, (float)b, 0.f, 1.f);
  for (int s = 0; s < NSYM; ++s) {
#pragma unroll
    for (int tx = 0; tx < NTX; ++tx) {
      __stcs(pH + tx * pitch2, v);
      __stcs(pL + tx * pitch2, v);
      __stcs(pM + tx * pitch2, v);
    }
    __stcs(pR, v);
    if (rx == 0) {
#pragma unroll
      for (int tx = 0; tx < NTX; ++tx) __stcs(pT + tx * pitch2, v);
    }
    pH += NRX * NTX * pitch2, pL += NRX * NTX * pitch2, pM += NRX * NTX * pitch2, pR += NRX * pitch2, pT += NTX * pitch2;
  }
}

// Mirror ownership with lane-pair exchange: lane t (f = t+1) computes bins +-f; even lanes store the +side pair
// (k1, k1+1), odd lanes the -side pair (k0, k0+1) as ONE 16-byte store.  `front` elements of padding in front of
// every row make the 256-byte chunks of a warp line-aligned when front = 4 and pitch = 608.
__global__ void __launch_bounds__(320, 2) mirror128_kernel(float2 *H, float2 *L, float2 *M, float2 *R, float2 *T, int pitch,
                                                           int front, int spin) {
  const int b = blockIdx.x / NRX, rx = blockIdx.x % NRX, t = threadIdx.x;
  const bool act = t < HALF;
  const int k = (t & 1) ? HALF - 1 - t : HALF + t;       // even start of this lane's pair
  const int64_t slot_h = (int64_t)NSYM * NRX * NTX * pitch, slot_r = (int64_t)NSYM * NRX * pitch,
                slot_t = (int64_t)NSYM * NTX * pitch;
  const int o = front + (act ? k : 0);
  float4 *pH = (float4 *)(H + b * slot_h + rx * NTX * pitch + o), *pL = (float4 *)(L + b * slot_h + rx * NTX * pitch + o),
         *pM = (float4 *)(M + b * slot_h + rx * NTX * pitch + o), *pR = (float4 *)(R + b * slot_r + rx * pitch + o),
         *pT = (float4 *)(T + b * slot_t + o);
  float4 v = make_float4((float)t, (float)b, 0.f, 1.f);
  const int p2 = pitch / 2;
  for (int s = 0; s < NSYM; ++s) {
    for (int i = 0; i < spin; ++i) v.x = fmaf(v.x, 1.0000001f, 1e-9f);
    if (act) {
#pragma unroll
      for (int tx = 0; tx < NTX; ++tx) {
        __stcs(pH + tx * p2, v);
        __stcs(pL + tx * p2, v);
        __stcs(pM + tx * p2, v);
      }
      __stcs(pR, v);
      if (rx == 0) {
#pragma unroll
        for (int tx = 0; tx < NTX; ++tx) __stcs(pT + tx * p2, v);
      }
    }
    pH += NRX * NTX * p2, pL += NRX * NTX * p2, pM += NRX * NTX * p2, pR += NRX * p2, pT += NTX * p2;
  }
}

// The product's 8-byte mirror stores, but with `front` elements of row padding (line-aligned chunks for front = 4, pitch 608)
__global__ void __launch_bounds__(320, 2) mirror64_front_kernel(float2 *H, float2 *L, float2 *M, float2 *R, float2 *T, int pitch,
                                                                int front, int spin) {
  const int b = blockIdx.x / NRX, rx = blockIdx.x % NRX, t = threadIdx.x;
  const bool v0 = t < HALF, v1 = t < HALF - 1;
  const int k0 = front + (v0 ? HALF - 1 - t : 0), k1 = front + (v1 ? HALF + t : 0);
  const int64_t slot_h = (int64_t)NSYM * NRX * NTX * pitch, slot_r = (int64_t)NSYM * NRX * pitch,
                slot_t = (int64_t)NSYM * NTX * pitch;
  float2 *pH = H + b * slot_h + rx * NTX * pitch, *pL = L + b * slot_h + rx * NTX * pitch,
         *pM = M + b * slot_h + rx * NTX * pitch, *pR = R + b * slot_r + rx * pitch, *pT = T + b * slot_t;
  float2 v = make_float2((float)t, (float)b);
  for (int s = 0; s < NSYM; ++s) {
    for (int i = 0; i < spin; ++i) v.x = fmaf(v.x, 1.0000001f, 1e-9f);
#pragma unroll
    for (int tx = 0; tx < NTX; ++tx) {
      if (v0) __stcs(pH + tx * pitch + k0, v);
      if (v1) __stcs(pH + tx * pitch + k1, v);
      if (v0) __stcs(pL + tx * pitch + k0, v);
      if (v1) __stcs(pL + tx * pitch + k1, v);
      if (v0) __stcs(pM + tx * pitch + k0, v);
      if (v1) __stcs(pM + tx * pitch + k1, v);
    }
    if (v0) __stcs(pR + k0, v);
    if (v1) __stcs(pR + k1, v);
    if (rx == 0) {
#pragma unroll
      for (int tx = 0; tx < NTX; ++tx) {
        if (v0) __stcs(pT + tx * pitch + k0, v);
        if (v1) __stcs(pT + tx * pitch + k1, v);
      }
    }
    pH += NRX * NTX * pitch, pL += NRX * NTX * pitch, pM += NRX * NTX * pitch, pR += NRX * pitch, pT += NTX * pitch;
  }
}

__global__ void fill_kernel(float4 *p, int64_t n) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) __stcs(p + i, v);
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main(int argc, char **argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 2048;
  const int pitch_max = 608;
  const size_t nH = (size_t)B * NSYM * NRX * NTX * pitch_max, nR = (size_t)B * NSYM * NRX * pitch_max;
  float2 *H, *L, *M, *R, *T;
  CK(cudaMalloc(&H, nH * 8 + 256));
  CK(cudaMalloc(&L, nH * 8 + 256));
  CK(cudaMalloc(&M, nH * 8 + 256));
  CK(cudaMalloc(&R, nR * 8 + 256));
  CK(cudaMalloc(&T, nR * 8 + 256));
  const double bytes = 8.0 * B * NSYM * NSC * (3.0 * NRX * NTX + NRX + NTX);     // algorithmic bytes (599 per row)
  printf("{\"B\": %d, \"algorithmic_GB\": %.3f", B, bytes / 1e9);
  {
    const int64_t n = (int64_t)(bytes / 16);
    float ms = time_ms([&] { fill_kernel<<<148 * 16, 256>>>((float4 *)H, n < (int64_t)(nH / 2) ? n : (int64_t)(nH / 2)); }, 5);
    const double fb = 16.0 * (n < (int64_t)(nH / 2) ? n : (int64_t)(nH / 2));
    printf(", \"fill_stcs_TBps\": %.3f", fb / ms / 1e9);
  }
#define RUN(name, ...)                                            \
  {                                                               \
    float ms = time_ms([&] { __VA_ARGS__; }, 5);                  \
    printf(", \"%s\": %.3f", name, bytes / ms / 1e9);             \
    fflush(stdout);                                               \
  }
  const dim3 g(B * NRX), blk(320);
  RUN("mirror_stcs_p599", pattern_kernel<0, 0><<<g, blk>>>(H, L, M, R, T, 599, 0))
  RUN("mirror_default_p599", pattern_kernel<1, 0><<<g, blk>>>(H, L, M, R, T, 599, 0))
  RUN("mirror_stwt_p599", pattern_kernel<2, 0><<<g, blk>>>(H, L, M, R, T, 599, 0))
  RUN("mirror_stcg_p599", pattern_kernel<3, 0><<<g, blk>>>(H, L, M, R, T, 599, 0))
  RUN("mirror_stcs_p600", pattern_kernel<0, 0><<<g, blk>>>(H, L, M, R, T, 600, 0))
  RUN("mirror_stcs_p608", pattern_kernel<0, 0><<<g, blk>>>(H, L, M, R, T, 608, 0))
  RUN("ascending_stcs_p599", pattern_kernel<0, 1><<<g, blk>>>(H, L, M, R, T, 599, 0))
  RUN("ascending_stcs_p600", pattern_kernel<0, 1><<<g, blk>>>(H, L, M, R, T, 600, 0))
  RUN("mirror_stcs_p599_spin64", pattern_kernel<0, 0><<<g, blk>>>(H, L, M, R, T, 599, 64))
  RUN("mirror_stcs_p599_spin256", pattern_kernel<0, 0><<<g, blk>>>(H, L, M, R, T, 599, 256))
  RUN("mirror_stcs_p599_spin512", pattern_kernel<0, 0><<<g, blk>>>(H, L, M, R, T, 599, 512))
  RUN("mirror_stcs_p599_spin1024", pattern_kernel<0, 0><<<g, blk>>>(H, L, M, R, T, 599, 1024))
  // occupancy: the product kernel runs 2 CTAs/SM (register-bound); dynamic smem pins the replay to 1, 2, 3 CTAs/SM
  CK(cudaFuncSetAttribute(pattern_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  RUN("mirror_stcs_p599_1cta", pattern_kernel<0, 0><<<g, blk, 200 * 1024>>>(H, L, M, R, T, 599, 0))
  RUN("mirror_stcs_p599_2cta", pattern_kernel<0, 0><<<g, blk, 100 * 1024>>>(H, L, M, R, T, 599, 0))
  RUN("mirror_stcs_p599_3cta", pattern_kernel<0, 0><<<g, blk, 70 * 1024>>>(H, L, M, R, T, 599, 0))
  RUN("mirror_stcs_p600_2cta", pattern_kernel<0, 0><<<g, blk, 100 * 1024>>>(H, L, M, R, T, 600, 0))
  RUN("mirror_stcs_p599_2cta_spin256", pattern_kernel<0, 0><<<g, blk, 100 * 1024>>>(H, L, M, R, T, 599, 256))
  RUN("mirror_stcs_p600_2cta_spin256", pattern_kernel<0, 0><<<g, blk, 100 * 1024>>>(H, L, M, R, T, 600, 256))
  CK(cudaFuncSetAttribute(pattern128_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  RUN("stg128_p600", pattern128_kernel<0><<<g, blk>>>((float4 *)H, (float4 *)L, (float4 *)M, (float4 *)R, (float4 *)T, 300))
  RUN("stg128_p600_2cta", pattern128_kernel<0><<<g, blk, 100 * 1024>>>((float4 *)H, (float4 *)L, (float4 *)M, (float4 *)R, (float4 *)T, 300))
  RUN("stg128_p608_2cta", pattern128_kernel<0><<<g, blk, 100 * 1024>>>((float4 *)H, (float4 *)L, (float4 *)M, (float4 *)R, (float4 *)T, 304))
  {
    const int64_t n8 = (int64_t)(bytes / 8) < (int64_t)nH ? (int64_t)(bytes / 8) : (int64_t)nH;
    float ms = time_ms([&] { fill8_kernel<<<148 * 16, 256>>>(H, n8); }, 5);
    printf(", \"fill8_stcs\": %.3f", 8.0 * n8 / ms / 1e9);
    const int ncta = B * NRX;
    const int64_t per = (int64_t)nH / ncta;      // ~ 14 * 4 rows per CTA, contiguous
    ms = time_ms([&] { private_region_kernel<float2><<<ncta, 320>>>(H, per); }, 5);
    printf(", \"private_region_f2\": %.3f", 8.0 * per * ncta / ms / 1e9);
    ms = time_ms([&] { private_region_kernel<float4><<<ncta, 320>>>((float4 *)H, per / 2); }, 5);
    printf(", \"private_region_f4\": %.3f", 8.0 * per * ncta / ms / 1e9);
    CK(cudaFuncSetAttribute(private_region_kernel<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ms = time_ms([&] { private_region_kernel<float2><<<ncta, 320, 100 * 1024>>>(H, per); }, 5);
    printf(", \"private_region_f2_2cta\": %.3f", 8.0 * per * ncta / ms / 1e9);
  }
  CK(cudaFuncSetAttribute(mirror128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(mirror64_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  RUN("mirror128_p608_f4_2cta", mirror128_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 608, 4, 0))
  RUN("mirror128_p608_f4_2cta_spin256", mirror128_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 608, 4, 256))
  RUN("mirror128_p608_f4_2cta_spin512", mirror128_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 608, 4, 512))
  RUN("mirror128_p608_f0_2cta", mirror128_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 608, 0, 0))
  RUN("mirror128_p600_f0_2cta", mirror128_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 600, 0, 0))
  RUN("mirror128_p600_f0_2cta_spin256", mirror128_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 600, 0, 256))
  RUN("mirror128_p604_f4_2cta", mirror128_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 604, 4, 0))
  RUN("mirror64_p608_f4_2cta", mirror64_front_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 608, 4, 0))
  RUN("mirror64_p608_f4_2cta_spin256", mirror64_front_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 608, 4, 256))
  RUN("mirror64_p608_f0_2cta", mirror64_front_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 608, 0, 0))
  RUN("mirror64_p600_f0_2cta", mirror64_front_kernel<<<g, blk, 100 * 1024>>>(H, L, M, R, T, 600, 0, 0))
  RUN("allrx_stcs_p599", pattern_allrx_kernel<0><<<B, blk>>>(H, L, M, R, T, 599))
  RUN("allrx_stcs_p600", pattern_allrx_kernel<0><<<B, blk>>>(H, L, M, R, T, 600))
  printf("}\n");
  return 0;
}
