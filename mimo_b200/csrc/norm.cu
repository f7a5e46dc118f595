// GroupNorm(+SiLU) and LayerNorm(+positional encoding) over channels-last activations. HBM-bound passes:
// 128-bit loads/stores, fp32 statistics, each thread owns a fixed 8-channel vector so gamma/beta/mean/rstd
// are loaded once and the loop over pixels is pure streaming.
#include <cuda_runtime.h>

#include "../../include/mimo_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace mimo {

// ------------------------------------------------------------------------------------------------
// GroupNorm
// ------------------------------------------------------------------------------------------------
struct GnArgs {
  const void* x0;
  const void* x1;
  const void* gamma;
  const void* beta;
  void* out;
  float* stats;  // [n][groups][2] = (sum, sumsq)
  int c0, c1, C, hw, groups, cpg;
  int vecs;      // C / 8
  int P;         // pixels processed per block iteration
  int pix_per_block;
  float eps;
  int silu;
};

template <bool kBf16>
__device__ __forceinline__ uint4 gn_load(const GnArgs& a, long long pix, int cv) {
  using C = Cvt<kBf16>;
  const int ch = cv * 8;
  if (ch < a.c0) {
    return *reinterpret_cast<const uint4*>(static_cast<const typename C::T*>(a.x0) + pix * a.c0 + ch);
  }
  return *reinterpret_cast<const uint4*>(static_cast<const typename C::T*>(a.x1) + pix * a.c1 + (ch - a.c0));
}

// pass 1: per-(image, group) sum and sum of squares
template <bool kBf16>
__global__ void __launch_bounds__(1024) gn_stats_kernel(GnArgs a) {
  using C = Cvt<kBf16>;
  __shared__ float s_sum[64], s_sq[64];
  const int n = blockIdx.y;
  const int cv = threadIdx.x % a.vecs;
  const int pl = threadIdx.x / a.vecs;
  const bool active = pl < a.P;
  for (int i = threadIdx.x; i < a.groups; i += blockDim.x) {
    s_sum[i] = 0.f;
    s_sq[i] = 0.f;
  }
  __syncthreads();
  const int ch0 = cv * 8;
  const int gA = ch0 / a.cpg;
  int split = (gA + 1) * a.cpg - ch0;  // channels [0, split) of this vector belong to gA, the rest to gA + 1
  if (split > 8) split = 8;
  float sA = 0.f, qA = 0.f, sB = 0.f, qB = 0.f;
  if (active) {
    const int p_begin = blockIdx.x * a.pix_per_block;
    int p_end = p_begin + a.pix_per_block;
    if (p_end > a.hw) p_end = a.hw;
    for (int p = p_begin + pl; p < p_end; p += a.P) {
      const uint4 u = gn_load<kBf16>(a, static_cast<long long>(n) * a.hw + p, cv);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = C::unpack(w[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < split) {
          sA += f[j];
          qA += f[j] * f[j];
        } else {
          sB += f[j];
          qB += f[j] * f[j];
        }
      }
    }
    atomicAdd(&s_sum[gA], sA);
    atomicAdd(&s_sq[gA], qA);
    if (split < 8) {
      atomicAdd(&s_sum[gA + 1], sB);
      atomicAdd(&s_sq[gA + 1], qB);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.groups; i += blockDim.x) {
    atomicAdd(&a.stats[(static_cast<long long>(n) * a.groups + i) * 2 + 0], s_sum[i]);
    atomicAdd(&a.stats[(static_cast<long long>(n) * a.groups + i) * 2 + 1], s_sq[i]);
  }
}

// pass 2: normalise + affine (+ SiLU), dense [n, hw, C] output (this is also where a virtual concat lands)
template <bool kBf16>
__global__ void __launch_bounds__(1024) gn_apply_kernel(GnArgs a) {
  using C = Cvt<kBf16>;
  const int n = blockIdx.y;
  const int cv = threadIdx.x % a.vecs;
  const int pl = threadIdx.x / a.vecs;
  if (pl >= a.P) return;
  const int ch0 = cv * 8;
  const float inv_cnt = 1.0f / (static_cast<float>(a.hw) * a.cpg);
  float sc[8], sh[8];
  {
    const uint4 ug = *reinterpret_cast<const uint4*>(static_cast<const typename C::T*>(a.gamma) + ch0);
    const uint4 ub = *reinterpret_cast<const uint4*>(static_cast<const typename C::T*>(a.beta) + ch0);
    const uint32_t wg[4] = {ug.x, ug.y, ug.z, ug.w};
    const uint32_t wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 g2 = C::unpack(wg[j]);
      const float2 b2 = C::unpack(wb[j]);
      sc[2 * j] = g2.x;
      sc[2 * j + 1] = g2.y;
      sh[2 * j] = b2.x;
      sh[2 * j + 1] = b2.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (ch0 + j) / a.cpg;
      const float s = a.stats[(static_cast<long long>(n) * a.groups + g) * 2 + 0];
      const float q = a.stats[(static_cast<long long>(n) * a.groups + g) * 2 + 1];
      const float mean = s * inv_cnt;
      float var = q * inv_cnt - mean * mean;
      var = var < 0.f ? 0.f : var;
      const float rstd = rsqrtf(var + a.eps);
      sh[j] = sh[j] - mean * rstd * sc[j];
      sc[j] = rstd * sc[j];
    }
  }
  const int p_begin = blockIdx.x * a.pix_per_block;
  int p_end = p_begin + a.pix_per_block;
  if (p_end > a.hw) p_end = a.hw;
  for (int p = p_begin + pl; p < p_end; p += a.P) {
    const long long pix = static_cast<long long>(n) * a.hw + p;
    const uint4 u = gn_load<kBf16>(a, pix, cv);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = C::unpack(w[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = f[j] * sc[j] + sh[j];
      if (a.silu) f[j] = silu_f(f[j]);
    }
    uint4 o;
    o.x = C::pack(f[0], f[1]);
    o.y = C::pack(f[2], f[3]);
    o.z = C::pack(f[4], f[5]);
    o.w = C::pack(f[6], f[7]);
    *reinterpret_cast<uint4*>(static_cast<typename C::T*>(a.out) + pix * a.C + ch0) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 2048)
// ------------------------------------------------------------------------------------------------
constexpr int kLnMaxVec = 8;  // uint4 per lane -> C <= 32 * 8 * 8 = 2048

template <bool kBf16>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x, const void* __restrict__ gamma, const void* __restrict__ beta,
                 void* __restrict__ out, long long rows, int Cdim, float eps, const void* __restrict__ pe,
                 long long rows_per_frame, int frames, int pe_off) {
  using C = Cvt<kBf16>;
  using T = typename C::T;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int vecs = Cdim >> 3;
  const T* xr = static_cast<const T*>(x) + row * Cdim;
  uint4 u[kLnMaxVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < vecs) {
      u[i] = *reinterpret_cast<const uint4*>(xr + v * 8);
      const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = C::unpack(w[j]);
        sum += t.x + t.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(Cdim);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < vecs) {
      const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = C::unpack(w[j]);
        const float d0 = t.x - mean, d1 = t.y - mean;
        sq += d0 * d0 + d1 * d1;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(Cdim) + eps);
  const T* per = nullptr;
  if (pe) per = static_cast<const T*>(pe) + (pe_off + (row / rows_per_frame) % frames) * Cdim;
  T* orow = static_cast<T*>(out) + row * Cdim;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < vecs) {
      const uint4 ug = __ldg(reinterpret_cast<const uint4*>(static_cast<const T*>(gamma) + v * 8));
      const uint4 ub = __ldg(reinterpret_cast<const uint4*>(static_cast<const T*>(beta) + v * 8));
      const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
      const uint32_t wg[4] = {ug.x, ug.y, ug.z, ug.w};
      const uint32_t wb[4] = {ub.x, ub.y, ub.z, ub.w};
      uint32_t wp[4] = {0, 0, 0, 0};
      if (per) {
        const uint4 up = __ldg(reinterpret_cast<const uint4*>(per + v * 8));
        wp[0] = up.x;
        wp[1] = up.y;
        wp[2] = up.z;
        wp[3] = up.w;
      }
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = C::unpack(w[j]);
        const float2 g2 = C::unpack(wg[j]);
        const float2 b2 = C::unpack(wb[j]);
        float y0 = (t.x - mean) * rstd * g2.x + b2.x;
        float y1 = (t.y - mean) * rstd * g2.y + b2.y;
        if (per) {
          // the reference rounds LN's output to the storage type before adding the encoding
          const float2 p2 = C::unpack(wp[j]);
          y0 = C::to_f(C::from_f(y0)) + p2.x;
          y1 = C::to_f(C::from_f(y1)) + p2.y;
        }
        o[j] = C::pack(y0, y1);
      }
      *reinterpret_cast<uint4*>(orow + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// C = 40 L channels (320 / 640 / 1280: every transformer width of the UNet), L = 8 / 16 / 32 lanes per row: each lane
// owns exactly five 8-channel vectors, a warp normalises 32 / L rows at a time and keeps gamma / beta (packed) in
// registers across kLnIter row groups. The generic kernel above spends ~450 instructions per 320-wide row (predicated
// 8-way unroll at 62 % lane use, gamma / beta re-read and unpacked per row, 64-bit frame arithmetic) and is issue-bound
// at 1.9 TB/s; this one needs ~100.
constexpr int kLnIter = 4;

template <bool kBf16, int L>
__global__ void __launch_bounds__(256, 2)
layernorm5_kernel(const void* __restrict__ x, const void* __restrict__ gamma, const void* __restrict__ beta,
                  void* __restrict__ out, int rows, float eps, const void* __restrict__ pe, int rows_per_frame,
                  int frames, int pe_off) {
  using C = Cvt<kBf16>;
  using T = typename C::T;
  constexpr int R = 32 / L;       // rows per warp pass
  constexpr int Cdim = 40 * L;
  const int lane = threadIdx.x & 31;
  const int sub = lane % L;
  const int rsel = lane / L;
  uint32_t g[5][4], b[5][4];  // packed pairs: unpacking on use is cheaper than 80 fp32 registers (occupancy)
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int v = sub + j * L;
    const uint4 ug = __ldg(reinterpret_cast<const uint4*>(static_cast<const T*>(gamma) + v * 8));
    const uint4 ub = __ldg(reinterpret_cast<const uint4*>(static_cast<const T*>(beta) + v * 8));
    g[j][0] = ug.x, g[j][1] = ug.y, g[j][2] = ug.z, g[j][3] = ug.w;
    b[j][0] = ub.x, b[j][1] = ub.y, b[j][2] = ub.z, b[j][3] = ub.w;
  }
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
#pragma unroll 1
  for (int it = 0; it < kLnIter; ++it) {
    const int row = (warp_global * kLnIter + it) * R + rsel;
    const bool ok = row < rows;
    const T* xr = static_cast<const T*>(x) + static_cast<long long>(row) * Cdim;
    uint4 u[5];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      u[j] = ok ? *reinterpret_cast<const uint4*>(xr + (sub + j * L) * 8) : make_uint4(0, 0, 0, 0);
      const uint32_t w[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 t = C::unpack(w[k]);
        sum += t.x + t.y;
      }
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / Cdim);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const uint32_t w[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 t = C::unpack(w[k]);
        const float d0 = t.x - mean, d1 = t.y - mean;
        sq += d0 * d0 + d1 * d1;
      }
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / Cdim) + eps);
    if (!ok) continue;
    const T* per = nullptr;
    if (pe) per = static_cast<const T*>(pe) + static_cast<long long>(pe_off + (row / rows_per_frame) % frames) * Cdim;
    T* orow = static_cast<T*>(out) + static_cast<long long>(row) * Cdim;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int v = sub + j * L;
      const uint32_t w[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
      uint32_t wp[4] = {0, 0, 0, 0};
      if (per) {
        const uint4 up = __ldg(reinterpret_cast<const uint4*>(per + v * 8));
        wp[0] = up.x;
        wp[1] = up.y;
        wp[2] = up.z;
        wp[3] = up.w;
      }
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 t = C::unpack(w[k]);
        const float2 g2 = C::unpack(g[j][k]);
        const float2 b2 = C::unpack(b[j][k]);
        float y0 = (t.x - mean) * rstd * g2.x + b2.x;
        float y1 = (t.y - mean) * rstd * g2.y + b2.y;
        if (per) {
          // the reference rounds LN's output to the storage type before adding the encoding
          const float2 p2 = C::unpack(wp[k]);
          y0 = C::to_f(C::from_f(y0)) + p2.x;
          y1 = C::to_f(C::from_f(y1)) + p2.y;
        }
        o[k] = C::pack(y0, y1);
      }
      *reinterpret_cast<uint4*>(orow + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <bool kBf16>
static void launch_ln5(int L, const void* x, const void* gamma, const void* beta, void* out, int rows, float eps,
                       const void* pe, int rpf, int frames, int pe_off, cudaStream_t st) {
  const int rows_per_block = 8 * kLnIter * (32 / L);
  const unsigned blocks = div_up(rows, rows_per_block);
  if (L == 8)
    layernorm5_kernel<kBf16, 8><<<blocks, 256, 0, st>>>(x, gamma, beta, out, rows, eps, pe, rpf, frames, pe_off);
  else if (L == 16)
    layernorm5_kernel<kBf16, 16><<<blocks, 256, 0, st>>>(x, gamma, beta, out, rows, eps, pe, rpf, frames, pe_off);
  else
    layernorm5_kernel<kBf16, 32><<<blocks, 256, 0, st>>>(x, gamma, beta, out, rows, eps, pe, rpf, frames, pe_off);
}

}  // namespace mimo

using namespace mimo;

extern "C" int mimo_groupnorm(const mimo_groupnorm_params* p, void* stream) {
  if (!p || !p->x0 || !p->gamma || !p->beta || !p->out || !p->stats)
    return set_error(MIMO_ERR_ARG, "mimo_groupnorm: null pointer");
  const int c1 = p->x1 ? p->c1 : 0;
  const int C = p->c0 + c1;
  if (p->n <= 0 || p->hw <= 0 || C <= 0 || p->groups <= 0 || p->groups > 64)
    return set_error(MIMO_ERR_ARG, "mimo_groupnorm: bad sizes");
  if ((p->c0 % 8) || (c1 % 8) || (C % p->groups) || (C / 8 > 1024))
    return set_error(MIMO_ERR_ARG, "mimo_groupnorm: channels must be multiples of 8 and divisible by groups");
  {
    const int cpg = C / p->groups;  // an 8-channel vector may straddle at most two groups
    if (!(cpg >= 8 || cpg == 4)) return set_error(MIMO_ERR_ARG, "mimo_groupnorm: channels per group must be 4 or >= 8");
  }
  if (int rc = ensure_device()) return rc;
  GnArgs a;
  a.x0 = p->x0;
  a.x1 = p->x1;
  a.gamma = p->gamma;
  a.beta = p->beta;
  a.out = p->out;
  a.stats = p->stats;
  a.c0 = p->c0;
  a.c1 = c1;
  a.C = C;
  a.hw = p->hw;
  a.groups = p->groups;
  a.cpg = C / p->groups;
  a.vecs = C / 8;
  a.P = 256 / a.vecs;
  if (a.P < 1) a.P = 1;
  if (a.P > p->hw) a.P = p->hw;
  int iters = 16;
  a.pix_per_block = a.P * iters;
  a.eps = p->eps;
  a.silu = p->silu;
  const int threads = ((a.vecs * a.P + 31) / 32) * 32;
  dim3 grid(div_up(p->hw, a.pix_per_block), p->n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(p->stats, 0, sizeof(float) * 2 * p->n * p->groups, st);
  if (e != cudaSuccess) return set_cuda_error("groupnorm memset", e);
  if (p->dtype == MIMO_BF16) {
    gn_stats_kernel<true><<<grid, threads, 0, st>>>(a);
    gn_apply_kernel<true><<<grid, threads, 0, st>>>(a);
  } else {
    gn_stats_kernel<false><<<grid, threads, 0, st>>>(a);
    gn_apply_kernel<false><<<grid, threads, 0, st>>>(a);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error("groupnorm launch", e);
  return MIMO_OK;
}

extern "C" int mimo_layernorm(const void* x, const void* gamma, const void* beta, void* out, int64_t rows,
                              int32_t c, float eps, const void* pe, int64_t rows_per_frame, int32_t frames,
                              int32_t pe_frame_offset, int32_t dtype, void* stream) {
  if (!x || !gamma || !beta || !out) return set_error(MIMO_ERR_ARG, "mimo_layernorm: null pointer");
  if (rows <= 0 || c <= 0 || (c % 8) || c > 32 * 8 * kLnMaxVec)
    return set_error(MIMO_ERR_ARG, "mimo_layernorm: c must be a multiple of 8 and <= 2048");
  if (pe && (rows_per_frame <= 0 || frames <= 0)) return set_error(MIMO_ERR_ARG, "mimo_layernorm: bad pe args");
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int L = c == 320 ? 8 : (c == 640 ? 16 : (c == 1280 ? 32 : 0));
  if (L && rows < (1LL << 31) && rows_per_frame < (1LL << 31)) {
    const int rpf = pe ? static_cast<int>(rows_per_frame) : 1;
    if (dtype == MIMO_BF16)
      launch_ln5<true>(L, x, gamma, beta, out, static_cast<int>(rows), eps, pe, rpf, frames, pe_frame_offset, st);
    else
      launch_ln5<false>(L, x, gamma, beta, out, static_cast<int>(rows), eps, pe, rpf, frames, pe_frame_offset, st);
    cudaError_t e5 = cudaGetLastError();
    if (e5 != cudaSuccess) return set_cuda_error("layernorm launch", e5);
    return MIMO_OK;
  }
  const unsigned blocks = div_up(rows, 8);
  if (dtype == MIMO_BF16)
    layernorm_kernel<true><<<blocks, 256, 0, st>>>(x, gamma, beta, out, rows, c, eps, pe, rows_per_frame, frames, pe_frame_offset);
  else
    layernorm_kernel<false><<<blocks, 256, 0, st>>>(x, gamma, beta, out, rows, c, eps, pe, rows_per_frame, frames, pe_frame_offset);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error("layernorm launch", e);
  return MIMO_OK;
}
