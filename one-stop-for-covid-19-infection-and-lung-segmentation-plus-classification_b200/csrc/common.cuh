// Shared device/host helpers for libb200unet (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/b200unet.h"

// ----------------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry returns int, message kept in a thread-local string
// ----------------------------------------------------------------------------------------------
void b2u_set_error(const char* fmt, ...);

#define B2U_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      b2u_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e),      \
                    cudaGetErrorString(_e));                                                  \
      return B2U_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define B2U_REQUIRE(cond, ...)                                                                \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      b2u_set_error(__VA_ARGS__);                                                             \
      return B2U_ERR_ARG;                                                                     \
    }                                                                                         \
  } while (0)

#define B2U_LAUNCH_CHECK() B2U_CHECK_CUDA(cudaGetLastError())

static inline int b2u_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static const int B2U_NUM_SMS = 148;

// ----------------------------------------------------------------------------------------------
// storage-type helpers. T is float (exact mode) or __half (tensor-core mode); math is always fp32.
// ----------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// 8 consecutive channels (C % 8 == 0, ld % 8 == 0 and 16B-aligned bases are required by callers)
template <typename T> __device__ __forceinline__ void load8(const T* p, float v[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float v[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__half>(const __half* p, float v[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
// 8 packed halves held in a register quad -> fp32
__device__ __forceinline__ void unpack8h(const uint4& u, float v[8]) {
  const __half2 h0 = *reinterpret_cast<const __half2*>(&u.x), h1 = *reinterpret_cast<const __half2*>(&u.y);
  const __half2 h2 = *reinterpret_cast<const __half2*>(&u.z), h3 = *reinterpret_cast<const __half2*>(&u.w);
  float2 f;
  f = __half22float2(h0); v[0] = f.x; v[1] = f.y;
  f = __half22float2(h1); v[2] = f.x; v[3] = f.y;
  f = __half22float2(h2); v[4] = f.x; v[5] = f.y;
  f = __half22float2(h3); v[6] = f.x; v[7] = f.y;
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float v[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float v[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__half>(__half* p, const float v[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// activations (Keras: relu, elu alpha=1).  dact takes the *output* y of the activation.
__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == B2U_ACT_RELU) return fmaxf(x, 0.f);
  if (act == B2U_ACT_ELU) return x > 0.f ? x : expm1f(x);
  return x;
}
__device__ __forceinline__ float act_bwd_from_y(float y, int act) {
  if (act == B2U_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == B2U_ACT_ELU) return y > 0.f ? 1.f : y + 1.f;
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-level per-channel partial sums of the streaming kernels.  Thread t owns the 8-channel group g = t % cg
// (cg = C / 8) and adds v[0..7] to sdst[g*8 .. g*8+7] in shared memory.  fp32 shared-memory atomicAdd compiles to a
// compare-and-swap retry loop on sm_100 (ATOMS.CAST.SPIN), so first fold the lanes of a warp that own the same group
// with shuffles (cg a power of two <= 16: lanes l and l ^ o share g for o >= cg) and issue one atomic per group and
// warp.  Must be called by all 32 lanes of the warp when cg is a power of two <= 16.
__device__ __forceinline__ void group_add8(float* sdst, const float v[8], int cg, int g) {
  if (cg <= 16 && (cg & (cg - 1)) == 0) {
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = v[k];
    for (int o = 16; o >= cg; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
    }
    if ((int)(threadIdx.x & 31) < cg) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&sdst[g * 8 + k], r[k]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&sdst[g * 8 + k], v[k]);
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ----------------------------------------------------------------------------------------------
// Philox4x32-10 dropout stream (restated on the CPU in oracle/philox.py)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint32_t dropout_threshold(float p) {
  return (uint32_t)floor((double)p * 4294967296.0);
}
// four keep-words for logical elements e4*4 .. e4*4+3
__device__ __forceinline__ uint4 dropout_words(uint64_t e4, uint64_t seed, uint64_t step, uint32_t op_id) {
  return philox4x32_10(make_uint4((uint32_t)e4, (uint32_t)step, op_id, (uint32_t)(e4 >> 32)),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
