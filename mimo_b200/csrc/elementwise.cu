// Layout converters, im2col gather, add / SiLU, and the fused CFG + DDIM update. All single coalesced passes.
#include <cuda_runtime.h>

#include "../../include/mimo_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace mimo {

// [b, c, f, h, w] -> [(b f), h, w, cpad]; one thread per output (pixel, 8-channel vector)
template <bool kBf16, typename SrcT>
__global__ void ncfhw_to_nhwc_kernel(const SrcT* __restrict__ src, void* __restrict__ dst, int b, int c, int f,
                                     int h, int w, int cpad) {
  using C = Cvt<kBf16>;
  const long long hw = static_cast<long long>(h) * w;
  const int vecs = cpad / 8;
  const long long total = static_cast<long long>(b) * f * hw * vecs;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % vecs);
    const long long pix = i / vecs;  // ((bi * f + fi) * hw + p)
    const long long p = pix % hw;
    const long long nf = pix / hw;
    const int fi = static_cast<int>(nf % f);
    const int bi = static_cast<int>(nf / f);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = cv * 8 + j;
      v[j] = 0.f;
      if (ch < c) {
        const long long s = ((static_cast<long long>(bi) * c + ch) * f + fi) * hw + p;
        if constexpr (sizeof(SrcT) == 4) {
          v[j] = src[s];
        } else {
          v[j] = C::to_f(reinterpret_cast<const typename C::T*>(src)[s]);
        }
      }
    }
    uint4 o;
    o.x = C::pack(v[0], v[1]);
    o.y = C::pack(v[2], v[3]);
    o.z = C::pack(v[4], v[5]);
    o.w = C::pack(v[6], v[7]);
    *reinterpret_cast<uint4*>(static_cast<typename C::T*>(dst) + pix * cpad + cv * 8) = o;
  }
}

// [(b f), h, w, ld] -> [b, c, f, h, w]; one thread per output element (w fastest -> coalesced writes)
template <bool kBf16, typename DstT>
__global__ void nhwc_to_ncfhw_kernel(const void* __restrict__ src, DstT* __restrict__ dst, int b, int c, int f,
                                     int h, int w, int ld) {
  using C = Cvt<kBf16>;
  const long long hw = static_cast<long long>(h) * w;
  const long long total = static_cast<long long>(b) * c * f * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i % hw;
    long long t = i / hw;
    const int fi = static_cast<int>(t % f);
    t /= f;
    const int ch = static_cast<int>(t % c);
    const int bi = static_cast<int>(t / c);
    const float v = C::to_f(static_cast<const typename C::T*>(src)[((static_cast<long long>(bi) * f + fi) * hw + p) * ld + ch]);
    if constexpr (sizeof(DstT) == 4) {
      dst[i] = v;
    } else {
      reinterpret_cast<typename C::T*>(dst)[i] = C::from_f(v);
    }
  }
}

// im2col for 3x3 windows: col[(n, oy, ox), tap * c + ch], optional stride and nearest-upsampled input.
// pad_lo is the number of zero rows/cols before the first input sample (1 for "padding=1"; 0 for the VAE
// encoder's asymmetric F.pad(0,1,0,1) + padding=0 downsample).
template <typename T>
__global__ void im2col3x3_kernel(const T* __restrict__ x, T* __restrict__ col, int n, int h, int w, int c,
                                 int stride, int upshift, int pad_lo, int oh, int ow, long long ldcol) {
  const int vecs = c / 8;
  const long long total = static_cast<long long>(n) * oh * ow * 9 * vecs;
  const int uh = h << upshift, uw = w << upshift;  // logical (upsampled) input size
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % vecs);
    long long t = i / vecs;
    const int tap = static_cast<int>(t % 9);
    const long long opix = t / 9;
    const int ox = static_cast<int>(opix % ow);
    const int oy = static_cast<int>((opix / ow) % oh);
    const long long ni = opix / (static_cast<long long>(ow) * oh);
    const int iy = oy * stride - pad_lo + tap / 3;
    const int ix = ox * stride - pad_lo + tap % 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < uh && ix >= 0 && ix < uw) {
      const long long sp = (ni * h + (iy >> upshift)) * w + (ix >> upshift);
      v = *reinterpret_cast<const uint4*>(x + sp * c + cv * 8);
    }
    *reinterpret_cast<uint4*>(col + opix * ldcol + static_cast<long long>(tap) * c + cv * 8) = v;
  }
}

// nearest x2 upsample, channels-last: one thread per output (pixel, 8-channel vector)
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int n, int h, int w, int vecs) {
  const int oh = 2 * h, ow = 2 * w;
  const long long total = static_cast<long long>(n) * oh * ow * vecs;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % vecs);
    const long long opix = i / vecs;
    const int ox = static_cast<int>(opix % ow);
    const int oy = static_cast<int>((opix / ow) % oh);
    const long long ni = opix / (static_cast<long long>(ow) * oh);
    out[i] = x[((ni * h + (oy >> 1)) * w + (ox >> 1)) * vecs + cv];
  }
}

// in-place row softmax, one CTA per row, fp32 math
template <bool kBf16>
__global__ void __launch_bounds__(256) softmax_rows_kernel(void* __restrict__ xp, int cols, long long ld) {
  using C = Cvt<kBf16>;
  using T = typename C::T;
  __shared__ float red[32];
  T* row = static_cast<T*>(xp) + static_cast<long long>(blockIdx.x) * ld;
  const int vecs = cols / 8;
  float m = -INFINITY;
  for (int v = threadIdx.x; v < vecs; v += blockDim.x) {
    const uint4 u = reinterpret_cast<const uint4*>(row)[v];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = C::unpack(w[j]);
      m = fmaxf(m, fmaxf(t.x, t.y));
    }
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int v = threadIdx.x; v < vecs; v += blockDim.x) {
    const uint4 u = reinterpret_cast<const uint4*>(row)[v];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = C::unpack(w[j]);
      s += __expf(t.x - m) + __expf(t.y - m);
    }
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
  const float inv = 1.0f / s;
  for (int v = threadIdx.x; v < vecs; v += blockDim.x) {
    const uint4 u = reinterpret_cast<const uint4*>(row)[v];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = C::unpack(w[j]);
      o[j] = C::pack(__expf(t.x - m) * inv, __expf(t.y - m) * inv);
    }
    reinterpret_cast<uint4*>(row)[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

template <bool kBf16, int kOp>  // 0: add, 1: silu, 2: quick-GELU x * sigmoid(1.702 x) (CLIP's hidden_act)
__global__ void ew_kernel(const void* __restrict__ a, const void* __restrict__ b, void* __restrict__ out,
                          long long nvec) {
  using C = Cvt<kBf16>;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 ua = static_cast<const uint4*>(a)[i];
    uint4 ub = make_uint4(0, 0, 0, 0);
    if (kOp == 0) ub = static_cast<const uint4*>(b)[i];
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w};
    const uint32_t wb[4] = {ub.x, ub.y, ub.z, ub.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = C::unpack(wa[j]);
      if (kOp == 0) {
        const float2 y = C::unpack(wb[j]);
        o[j] = C::pack(x.x + y.x, x.y + y.y);
      } else if (kOp == 1) {
        o[j] = C::pack(silu_f(x.x), silu_f(x.y));
      } else {
        o[j] = C::pack(__fdividef(x.x, 1.0f + __expf(-1.702f * x.x)), __fdividef(x.y, 1.0f + __expf(-1.702f * x.y)));
      }
    }
    static_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// CFG + DDIM (v-prediction, eta = 0). Every intermediate is rounded to the storage type where the reference's
// torch expression would round it (it runs the whole update in the latents' dtype).
template <bool kBf16>
__global__ void cfg_ddim_kernel(const void* __restrict__ pu, const void* __restrict__ pc,
                                const void* __restrict__ counter, long long frame_stride, int frames,
                                void* __restrict__ lat, long long count, float g, float sa_t, float s1a_t,
                                float sa_p, float s1a_p) {
  using C = Cvt<kBf16>;
  using T = typename C::T;
  auto rnd = [](float v) { return C::to_f(C::from_f(v)); };
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float u = C::to_f(static_cast<const T*>(pu)[i]);
    float c = C::to_f(static_cast<const T*>(pc)[i]);
    if (counter) {
      const float cnt = C::to_f(static_cast<const T*>(counter)[(i / frame_stride) % frames]);
      u = rnd(u / cnt);
      c = rnd(c / cnt);
    }
    const float v = rnd(u + rnd(g * rnd(c - u)));
    const float x = C::to_f(static_cast<T*>(lat)[i]);
    // DDIMScheduler.step: alpha terms are fp32 scalars -> products promote to fp32 only for 0-dim tensors'
    // python floats; tensor math stays in the storage dtype
    const float x0 = rnd(rnd(sa_t * x) - rnd(s1a_t * v));
    const float e = rnd(rnd(sa_t * v) + rnd(s1a_t * x));
    const float dir = rnd(s1a_p * e);
    const float prev = rnd(rnd(sa_p * x0) + dir);
    static_cast<T*>(lat)[i] = C::from_f(prev);
  }
}

static inline unsigned ew_grid(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace mimo

using namespace mimo;

#define MIMO_CHECK_LAUNCH(what)                                  \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return set_cuda_error(what, e__);    \
  } while (0)

extern "C" int mimo_ncfhw_to_nhwc(const void* src, void* dst, int32_t b, int32_t c, int32_t f, int32_t h,
                                  int32_t w, int32_t cpad, int32_t src_is_f32, int32_t dtype, void* stream) {
  if (!src || !dst || b <= 0 || c <= 0 || f <= 0 || h <= 0 || w <= 0 || cpad < c || (cpad % 8))
    return set_error(MIMO_ERR_ARG, "mimo_ncfhw_to_nhwc: bad arguments");
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(b) * f * h * w * (cpad / 8);
  const unsigned grid = ew_grid(total, 256);
  if (src_is_f32) {
    if (dtype == MIMO_BF16)
      ncfhw_to_nhwc_kernel<true, float><<<grid, 256, 0, st>>>(static_cast<const float*>(src), dst, b, c, f, h, w, cpad);
    else
      ncfhw_to_nhwc_kernel<false, float><<<grid, 256, 0, st>>>(static_cast<const float*>(src), dst, b, c, f, h, w, cpad);
  } else {
    if (dtype == MIMO_BF16)
      ncfhw_to_nhwc_kernel<true, uint16_t><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(src), dst, b, c, f, h, w, cpad);
    else
      ncfhw_to_nhwc_kernel<false, uint16_t><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(src), dst, b, c, f, h, w, cpad);
  }
  MIMO_CHECK_LAUNCH("ncfhw_to_nhwc launch");
  return MIMO_OK;
}

extern "C" int mimo_nhwc_to_ncfhw(const void* src, void* dst, int32_t b, int32_t c, int32_t f, int32_t h,
                                  int32_t w, int32_t ld, int32_t dst_is_f32, int32_t dtype, void* stream) {
  if (!src || !dst || b <= 0 || c <= 0 || f <= 0 || h <= 0 || w <= 0 || ld < c)
    return set_error(MIMO_ERR_ARG, "mimo_nhwc_to_ncfhw: bad arguments");
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(b) * c * f * h * w;
  const unsigned grid = ew_grid(total, 256);
  if (dst_is_f32) {
    if (dtype == MIMO_BF16)
      nhwc_to_ncfhw_kernel<true, float><<<grid, 256, 0, st>>>(src, static_cast<float*>(dst), b, c, f, h, w, ld);
    else
      nhwc_to_ncfhw_kernel<false, float><<<grid, 256, 0, st>>>(src, static_cast<float*>(dst), b, c, f, h, w, ld);
  } else {
    if (dtype == MIMO_BF16)
      nhwc_to_ncfhw_kernel<true, uint16_t><<<grid, 256, 0, st>>>(src, static_cast<uint16_t*>(dst), b, c, f, h, w, ld);
    else
      nhwc_to_ncfhw_kernel<false, uint16_t><<<grid, 256, 0, st>>>(src, static_cast<uint16_t*>(dst), b, c, f, h, w, ld);
  }
  MIMO_CHECK_LAUNCH("nhwc_to_ncfhw launch");
  return MIMO_OK;
}

extern "C" int mimo_im2col3x3(const void* x, void* col, int32_t n, int32_t h, int32_t w, int32_t c,
                              int32_t stride, int32_t upshift, int32_t pad_lo, int64_t ldcol, int32_t dtype,
                              void* stream) {
  (void)dtype;  // 16-bit payload either way
  if (!x || !col || n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8) || stride < 1 || stride > 2 || upshift < 0 ||
      upshift > 1 || pad_lo < 0 || pad_lo > 1 || ldcol < 9LL * c || (ldcol % 8))
    return set_error(MIMO_ERR_ARG, "mimo_im2col3x3: bad arguments");
  if (int rc = ensure_device()) return rc;
  const int uh = h << upshift, uw = w << upshift;
  // output size of a 3x3 window: pad_lo = 1 -> "padding 1"; pad_lo = 0 -> input padded by one at the far edge
  const int oh = (uh + 2 * pad_lo - 3 + (pad_lo ? 0 : 1)) / stride + 1;
  const int ow = (uw + 2 * pad_lo - 3 + (pad_lo ? 0 : 1)) / stride + 1;
  const long long total = static_cast<long long>(n) * oh * ow * 9 * (c / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  im2col3x3_kernel<uint16_t><<<ew_grid(total, 256), 256, 0, st>>>(
      static_cast<const uint16_t*>(x), static_cast<uint16_t*>(col), n, h, w, c, stride, upshift, pad_lo, oh, ow,
      ldcol);
  MIMO_CHECK_LAUNCH("im2col3x3 launch");
  return MIMO_OK;
}

extern "C" int mimo_upsample2x(const void* x, void* out, int32_t n, int32_t h, int32_t w, int32_t c, int32_t dtype,
                               void* stream) {
  (void)dtype;
  if (!x || !out || n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8)) return set_error(MIMO_ERR_ARG, "mimo_upsample2x: bad arguments");
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(n) * 4 * h * w * (c / 8);
  upsample2x_kernel<<<ew_grid(total, 256), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<uint4*>(out), n, h, w, c / 8);
  MIMO_CHECK_LAUNCH("upsample2x launch");
  return MIMO_OK;
}

extern "C" int mimo_softmax_rows(void* x, int64_t rows, int32_t cols, int64_t ld, int32_t dtype, void* stream) {
  if (!x || rows <= 0 || cols <= 0 || (cols % 8) || (ld % 8) || rows > 0x7fffffffLL)
    return set_error(MIMO_ERR_ARG, "mimo_softmax_rows: bad arguments");
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == MIMO_BF16)
    softmax_rows_kernel<true><<<static_cast<unsigned>(rows), 256, 0, st>>>(x, cols, ld);
  else
    softmax_rows_kernel<false><<<static_cast<unsigned>(rows), 256, 0, st>>>(x, cols, ld);
  MIMO_CHECK_LAUNCH("softmax_rows launch");
  return MIMO_OK;
}

extern "C" int mimo_add(const void* a, const void* b, void* out, int64_t count, int32_t dtype, void* stream) {
  if (!a || !b || !out || count <= 0 || (count % 8)) return set_error(MIMO_ERR_ARG, "mimo_add: bad arguments");
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nvec = count / 8;
  if (dtype == MIMO_BF16)
    ew_kernel<true, 0><<<ew_grid(nvec, 256), 256, 0, st>>>(a, b, out, nvec);
  else
    ew_kernel<false, 0><<<ew_grid(nvec, 256), 256, 0, st>>>(a, b, out, nvec);
  MIMO_CHECK_LAUNCH("add launch");
  return MIMO_OK;
}

extern "C" int mimo_silu(const void* x, void* out, int64_t count, int32_t dtype, void* stream) {
  if (!x || !out || count <= 0 || (count % 8)) return set_error(MIMO_ERR_ARG, "mimo_silu: bad arguments");
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nvec = count / 8;
  if (dtype == MIMO_BF16)
    ew_kernel<true, 1><<<ew_grid(nvec, 256), 256, 0, st>>>(x, nullptr, out, nvec);
  else
    ew_kernel<false, 1><<<ew_grid(nvec, 256), 256, 0, st>>>(x, nullptr, out, nvec);
  MIMO_CHECK_LAUNCH("silu launch");
  return MIMO_OK;
}

extern "C" int mimo_quick_gelu(const void* x, void* out, int64_t count, int32_t dtype, void* stream) {
  if (!x || !out || count <= 0 || (count % 8)) return set_error(MIMO_ERR_ARG, "mimo_quick_gelu: bad arguments");
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nvec = count / 8;
  if (dtype == MIMO_BF16)
    ew_kernel<true, 2><<<ew_grid(nvec, 256), 256, 0, st>>>(x, nullptr, out, nvec);
  else
    ew_kernel<false, 2><<<ew_grid(nvec, 256), 256, 0, st>>>(x, nullptr, out, nvec);
  MIMO_CHECK_LAUNCH("quick_gelu launch");
  return MIMO_OK;
}

extern "C" int mimo_cfg_ddim_step(const void* pred_uncond, const void* pred_cond, const void* counter_or_null,
                                  int64_t frame_stride, void* latents, int64_t count, float guidance,
                                  float sqrt_a_t, float sqrt_1ma_t, float sqrt_a_prev, float sqrt_1ma_prev,
                                  int32_t dtype, void* stream) {
  if (!pred_uncond || !pred_cond || !latents || count <= 0)
    return set_error(MIMO_ERR_ARG, "mimo_cfg_ddim_step: bad arguments");
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // counter (if given) holds one value per frame; latents are [1, 4, F, h, w] so the frame index of element i is
  // (i / frame_stride) % frames with frames = count / (4 * frame_stride)
  int frames = 1;
  if (counter_or_null) {
    if (frame_stride <= 0 || count % (4 * frame_stride)) return set_error(MIMO_ERR_ARG, "mimo_cfg_ddim_step: bad frame_stride");
    frames = static_cast<int>(count / (4 * frame_stride));
  }
  if (dtype == MIMO_BF16)
    cfg_ddim_kernel<true><<<ew_grid(count, 256), 256, 0, st>>>(pred_uncond, pred_cond, counter_or_null, frame_stride,
                                                              frames, latents, count, guidance, sqrt_a_t,
                                                              sqrt_1ma_t, sqrt_a_prev, sqrt_1ma_prev);
  else
    cfg_ddim_kernel<false><<<ew_grid(count, 256), 256, 0, st>>>(pred_uncond, pred_cond, counter_or_null,
                                                               frame_stride, frames, latents, count, guidance,
                                                               sqrt_a_t, sqrt_1ma_t, sqrt_a_prev, sqrt_1ma_prev);
  MIMO_CHECK_LAUNCH("cfg_ddim launch");
  return MIMO_OK;
}
