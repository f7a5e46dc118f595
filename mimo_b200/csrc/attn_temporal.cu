// Motion-module temporal self-attention: for every (batch b, pixel p, head h) an Fq x F attention over the
// frame axis (F <= 32, head dim d a multiple of 8).
//
// One CTA per (b, p): its 8 (= heads) warps stage the K and V rows of all F frames of that pixel in shared
// memory with 128-bit loads (each row is C contiguous elements), then warp h / lane j computes query row j of
// head h: scores against the F keys (K rows are smem broadcasts), fp32 softmax, and the PV product 8 output
// channels at a time. The "(b f) d c <-> (b d) f c" transposes of the reference never materialise: frames are
// addressed by stride. The op is ~0.2 % of the UNet's FLOPs and HBM/latency bound, hence CUDA cores.
//
// Frame sharding: with the clip's frames split over G GPUs the queries are the Fq local frames while K/V cover all
// F frames, stored as G chunks of frames_per_chunk frames each (the all-gathered per-rank buffers):
//   kv_row(b, f, p) = (f / fpc) * chunk_stride_rows + (b * fpc + f % fpc) * hw + p
#include <cuda_runtime.h>

#include "../../include/mimo_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace mimo {

struct TemporalArgs {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  long long ld_q, ld_kv, ld_out, chunk_stride_rows;
  int Fq, F, fpc, hw, heads, d;
  float scale;
};

template <bool kBf16>
__global__ void __launch_bounds__(1024) attn_temporal_kernel(TemporalArgs a) {
  using C = Cvt<kBf16>;
  using T = typename C::T;
  extern __shared__ uint4 smem_kv[];  // K[F][Cdim] then V[F][Cdim], 16-bit
  const int Cdim = a.heads * a.d;
  const int vecs = Cdim / 8;
  T* sk = reinterpret_cast<T*>(smem_kv);
  T* sv = sk + static_cast<size_t>(a.F) * Cdim;
  const int b = blockIdx.x / a.hw;
  const int p = blockIdx.x % a.hw;

  for (int i = threadIdx.x; i < a.F * vecs; i += blockDim.x) {
    const int f = i / vecs, cv = i % vecs;
    const long long row = static_cast<long long>(f / a.fpc) * a.chunk_stride_rows +
                          (static_cast<long long>(b) * a.fpc + f % a.fpc) * a.hw + p;
    const long long off = row * a.ld_kv + cv * 8;
    reinterpret_cast<uint4*>(sk)[i] = *reinterpret_cast<const uint4*>(static_cast<const T*>(a.k) + off);
    reinterpret_cast<uint4*>(sv)[i] = *reinterpret_cast<const uint4*>(static_cast<const T*>(a.v) + off);
  }
  __syncthreads();

  const int h = threadIdx.x >> 5;
  const int j = threadIdx.x & 31;
  if (h >= a.heads || j >= a.Fq) return;
  const int F = a.F, d = a.d;
  const long long qtok = (static_cast<long long>(b) * a.Fq + j) * a.hw + p;

  const T* qrow = static_cast<const T*>(a.q) + qtok * a.ld_q + h * d;
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
      s[i] *= a.scale;
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
  T* orow = static_cast<T*>(a.out) + qtok * a.ld_out + h * d;
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

extern "C" int mimo_attn_temporal(const mimo_attn_temporal_params* p, void* stream) {
  if (!p || !p->q || !p->k || !p->v || !p->out) return set_error(MIMO_ERR_ARG, "mimo_attn_temporal: null pointer");
  const int fpc = p->frames_per_chunk > 0 ? p->frames_per_chunk : p->kv_frames;
  if (p->batch <= 0 || p->q_frames <= 0 || p->kv_frames <= 0 || p->kv_frames > 32 || p->q_frames > 32 || p->hw <= 0 ||
      p->heads <= 0 || p->heads > 32 || p->d <= 0 || (p->d % 8) || (p->ld_q % 8) || (p->ld_kv % 8) || (p->ld_out % 8) ||
      (p->kv_frames % fpc))
    return set_error(MIMO_ERR_ARG, "mimo_attn_temporal: need frames <= 32, heads <= 32, d % 8 == 0, kv_frames % chunk == 0");
  if (int rc = ensure_device()) return rc;
  const size_t smem = static_cast<size_t>(2) * p->kv_frames * p->heads * p->d * 2;
  if (smem > 227 * 1024) return set_error(MIMO_ERR_ARG, "mimo_attn_temporal: K/V tile exceeds shared memory");
  TemporalArgs a;
  a.q = p->q;
  a.k = p->k;
  a.v = p->v;
  a.out = p->out;
  a.ld_q = p->ld_q;
  a.ld_kv = p->ld_kv;
  a.ld_out = p->ld_out;
  a.chunk_stride_rows = p->chunk_stride_rows;
  a.Fq = p->q_frames;
  a.F = p->kv_frames;
  a.fpc = fpc;
  a.hw = p->hw;
  a.heads = p->heads;
  a.d = p->d;
  a.scale = p->scale;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = p->heads * 32;
  const unsigned grid = static_cast<unsigned>(p->batch) * p->hw;
  static bool attr_done[2] = {false, false};
  cudaError_t e;
  if (p->dtype == MIMO_BF16) {
    if (!attr_done[1]) {
      e = cudaFuncSetAttribute(attn_temporal_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return set_cuda_error("attn_temporal attr", e);
      attr_done[1] = true;
    }
    attn_temporal_kernel<true><<<grid, threads, smem, st>>>(a);
  } else {
    if (!attr_done[0]) {
      e = cudaFuncSetAttribute(attn_temporal_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return set_cuda_error("attn_temporal attr", e);
      attr_done[0] = true;
    }
    attn_temporal_kernel<false><<<grid, threads, smem, st>>>(a);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error("attn_temporal launch", e);
  return MIMO_OK;
}
