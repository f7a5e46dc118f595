// Motion-module temporal self-attention: for every (batch b, pixel p, head h) an F x F attention over the
// frame axis (F <= 32, head dim d in {40, 80, 160, ...} multiple of 8).
//
// One CTA per (b, p): its 8 (= heads) warps stage the K and V rows of all F frames of that pixel in shared
// memory with 128-bit loads (each row is C contiguous elements), then warp h / lane j computes query row j of
// head h: scores against the F keys (K rows are smem broadcasts), fp32 softmax, and the PV product 8 output
// channels at a time. The "(b f) d c <-> (b d) f c" transposes of the reference never materialise: frames are
// addressed by stride. The op is ~0.2 % of the UNet's FLOPs and HBM/latency bound, hence CUDA cores.
#include <cuda_runtime.h>

#include "../../include/mimo_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace mimo {

template <bool kBf16>
__global__ void __launch_bounds__(1024)
attn_temporal_kernel(const void* __restrict__ qp, const void* __restrict__ kp, const void* __restrict__ vp,
                     long long ld, void* __restrict__ op, long long ldo, int F, int hw, int heads, int d,
                     float scale) {
  using C = Cvt<kBf16>;
  using T = typename C::T;
  extern __shared__ uint4 smem_kv[];  // K[F][Cdim] then V[F][Cdim], 16-bit
  const int Cdim = heads * d;
  const int vecs = Cdim / 8;
  T* sk = reinterpret_cast<T*>(smem_kv);
  T* sv = sk + static_cast<size_t>(F) * Cdim;
  const int b = blockIdx.x / hw;
  const int p = blockIdx.x % hw;
  const long long tok0 = static_cast<long long>(b) * F * hw + p;  // token of frame 0; frame f adds f * hw

  for (int i = threadIdx.x; i < F * vecs; i += blockDim.x) {
    const int f = i / vecs, cv = i % vecs;
    const long long off = (tok0 + static_cast<long long>(f) * hw) * ld + cv * 8;
    reinterpret_cast<uint4*>(sk)[i] = *reinterpret_cast<const uint4*>(static_cast<const T*>(kp) + off);
    reinterpret_cast<uint4*>(sv)[i] = *reinterpret_cast<const uint4*>(static_cast<const T*>(vp) + off);
  }
  __syncthreads();

  const int h = threadIdx.x >> 5;
  const int j = threadIdx.x & 31;
  if (h >= heads || j >= F) return;

  const T* qrow = static_cast<const T*>(qp) + (tok0 + static_cast<long long>(j) * hw) * ld + h * d;
  float s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = 0.f;
  for (int c8 = 0; c8 < d / 8; ++c8) {
    const uint4 uq = *reinterpret_cast<const uint4*>(qrow + c8 * 8);
    const uint32_t wq[4] = {uq.x, uq.y, uq.z, uq.w};
    float q[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 v2 = C::unpack(wq[t]);
      q[2 * t] = v2.x;
      q[2 * t + 1] = v2.y;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i < F) {
        const uint4 uk = *reinterpret_cast<const uint4*>(sk + static_cast<size_t>(i) * Cdim + h * d + c8 * 8);
        const uint32_t wk[4] = {uk.x, uk.y, uk.z, uk.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 v2 = C::unpack(wk[t]);
          s[i] = fmaf(q[2 * t], v2.x, s[i]);
          s[i] = fmaf(q[2 * t + 1], v2.y, s[i]);
        }
      }
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < F) {
      s[i] *= scale;
      m = fmaxf(m, s[i]);
    }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < F) {
      s[i] = __expf(s[i] - m);
      sum += s[i];
    }
  const float inv = 1.0f / sum;
  T* orow = static_cast<T*>(op) + (tok0 + static_cast<long long>(j) * hw) * ldo + h * d;
  for (int c8 = 0; c8 < d / 8; ++c8) {
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i < F) {
        const uint4 uv = *reinterpret_cast<const uint4*>(sv + static_cast<size_t>(i) * Cdim + h * d + c8 * 8);
        const uint32_t wv[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 v2 = C::unpack(wv[t]);
          o[2 * t] = fmaf(s[i], v2.x, o[2 * t]);
          o[2 * t + 1] = fmaf(s[i], v2.y, o[2 * t + 1]);
        }
      }
    }
    uint4 w;
    w.x = C::pack(o[0] * inv, o[1] * inv);
    w.y = C::pack(o[2] * inv, o[3] * inv);
    w.z = C::pack(o[4] * inv, o[5] * inv);
    w.w = C::pack(o[6] * inv, o[7] * inv);
    *reinterpret_cast<uint4*>(orow + c8 * 8) = w;
  }
}

}  // namespace mimo

using namespace mimo;

extern "C" int mimo_attn_temporal(const void* q, const void* k, const void* v, int64_t ld_qkv, void* out,
                                  int64_t ld_out, int32_t batch, int32_t frames, int32_t hw, int32_t heads,
                                  int32_t d, float scale, int32_t dtype, void* stream) {
  if (!q || !k || !v || !out) return set_error(MIMO_ERR_ARG, "mimo_attn_temporal: null pointer");
  if (batch <= 0 || frames <= 0 || frames > 32 || hw <= 0 || heads <= 0 || heads > 32 || d <= 0 || (d % 8) ||
      (ld_qkv % 8) || (ld_out % 8))
    return set_error(MIMO_ERR_ARG, "mimo_attn_temporal: need frames <= 32, heads <= 32, d % 8 == 0");
  if (int rc = ensure_device()) return rc;
  const size_t smem = static_cast<size_t>(2) * frames * heads * d * 2;
  if (smem > 227 * 1024) return set_error(MIMO_ERR_ARG, "mimo_attn_temporal: K/V tile exceeds shared memory");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = heads * 32;
  const unsigned grid = static_cast<unsigned>(batch) * hw;
  cudaError_t e;
  if (dtype == MIMO_BF16) {
    e = cudaFuncSetAttribute(attn_temporal_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_cuda_error("attn_temporal attr", e);
    attn_temporal_kernel<true><<<grid, threads, smem, st>>>(q, k, v, ld_qkv, out, ld_out, frames, hw, heads, d, scale);
  } else {
    e = cudaFuncSetAttribute(attn_temporal_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_cuda_error("attn_temporal attr", e);
    attn_temporal_kernel<false><<<grid, threads, smem, st>>>(q, k, v, ld_qkv, out, ld_out, frames, hw, heads, d, scale);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error("attn_temporal launch", e);
  return MIMO_OK;
}
