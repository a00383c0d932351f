// Philox4x32-10 counter-based generator (Salmon et al., SC'11) and the draw layout of b2c.h.
// oracle/philox.py is the bit-exact CPU twin used by the parity tests.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2c {

enum : uint32_t { STREAM_SYMBOLS = 0, STREAM_JAKES = 1, STREAM_NOISE = 2, STREAM_PARAMS = 3 };

struct PhiloxKey {
  uint32_t k0, k1;   // seed
  uint32_t s0, s1;   // global slot index (counter words 2, 3)
};

__host__ __device__ __forceinline__ PhiloxKey make_key(uint64_t seed, int64_t slot) {
  PhiloxKey k;
  k.k0 = (uint32_t)seed;
  k.k1 = (uint32_t)(seed >> 32);
  k.s0 = (uint32_t)(uint64_t)slot;
  k.s1 = (uint32_t)((uint64_t)slot >> 32);
  return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 draw(const PhiloxKey &k, uint32_t stream, uint32_t index) {
  return philox4x32_10(index, stream, k.s0, k.s1, k.k0, k.k1);
}

// 23-bit uniform u = (m + 1/2) 2^-23 in (0,1), m = the word's top 23 bits.  Built from the bits of 1 + m 2^-23 (one
// shift, one OR) instead of an int -> float conversion and a multiply; every step is exact in fp32 (the result has 24
// significant bits), so the CPU twin matches bit for bit.
__device__ __forceinline__ float one_plus_m(uint32_t w) { return __uint_as_float(0x3f800000u | (w >> 9)); }
__device__ __forceinline__ float u01(uint32_t w) { return one_plus_m(w) - 0.99999994039535522f; }   // 1 - 2^-24

// 2 pi (u - 1/2) for the same u: (1 + m 2^-23) - 3/2 is exact, the half-step 2^-24 rides in the FMA's addend.
__device__ __forceinline__ float two_pi_u_centered(uint32_t w) {
  return fmaf(6.283185307179586f, one_plus_m(w) - 1.5f, 6.283185307179586f * 5.9604644775390625e-08f);
}

// exp(j 2 pi u) for the uniform of word w.  The SFU wants its argument in [-pi, pi]: flipping the top mantissa bit
// of m gives u' = u +- 1/2 (mod 1), and 2 pi (u' - 1/2) is the same angle in (-pi, pi) -- no round / subtract.
__device__ __forceinline__ float2 cis_u01(uint32_t w) {
  float s, c;
  __sincosf(two_pi_u_centered(w ^ 0x80000000u), &s, &c);
  return make_float2(c, s);
}

// Box-Muller pair -> one complex normal with unit variance per component.
__device__ __forceinline__ float2 normal_pair(uint32_t w1, uint32_t w2) {
  // r = sqrt(-2 ln u) on the SFU: lg2.approx (abs. error <= 2^-22 on [0.5, 2]) and rsqrt.approx; the
  // clamp keeps r finite should the approximation round ln u to +0 for u within 1e-7 of 1.
  float x = fmaxf(-1.3862943611198906f * __log2f(u01(w1)), 1e-30f);
  float r = x * rsqrtf(x);
  float s, c;                         // cos(2 pi u) = -cos(2 pi (u - 1/2)), same for sin
  __sincosf(two_pi_u_centered(w2), &s, &c);
  return make_float2(-r * c, -r * s);
}

__device__ __forceinline__ uint32_t pick(uint4 w, int i) {
  return i == 0 ? w.x : (i == 1 ? w.y : (i == 2 ? w.z : w.w));
}

}  // namespace b2c
