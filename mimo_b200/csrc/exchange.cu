// Frame-shard <-> pixel-shard exchange of the motion module over NVLink peer memory (no NCCL on the data path).
//
// A clip's frames are sharded over the G GPUs of a frame group; everything in the UNet is per frame except the
// motion module's attention over the frame axis (src/models/motion_module.py:353-390). Around every motion module the
// tokens are re-sharded: frames -> pixels (each GPU then holds ALL frames of hw / G pixels, so the whole temporal
// transformer block is local) and back. Each GPU PULLS its share straight out of its peers' HBM (ld.global on
// IPC-mapped peer pointers; ~775 GB/s per direction through NVSwitch) and writes locally, fused with the residual add
// on the way back. No packing pass, no collective library, no host involvement: the kernel is an ordinary node of the
// captured CUDA graph.
//
// Protocol (per frame group; `e` = 1, 2, 3, ... counts exchanges, kept on the device so that graph replays advance it):
//   1. announce: write e into ready[r] of every peer          ("my source buffer holds exchange e")
//   2. wait until my own ready[s] >= e for every peer s        (peer data is complete and visible)
//   3. pull: peer loads -> local stores (+ residual)
//   4. the last block to finish publishes epoch e + 1 for the next launch.
// Buffer reuse needs no second flag: a peer that announced e has finished (stream order) its exchange e - 1, i.e. all
// its reads of my previous source buffer; and the producer that overwrites a source buffer runs, in stream order, after
// the exchange in between (two source buffers: one per direction). Every spin is bounded: a diverged rank traps the
// kernel (cudaErrorLaunchFailure) instead of hanging the GPU.
#include <cuda_runtime.h>

#include "../../include/mimo_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace mimo {

struct XchgArgs {
  const void* src[MIMO_MAX_PEERS];
  unsigned* ready[MIMO_MAX_PEERS];
  unsigned* ctl;  // [0] epoch of the next exchange (>= 1), [1] finished-block ticket
  void* dst;
  const void* residual;
  int mode, G, r, b, fl, hw, C;
  int seg_vecs;  // 16-byte vectors per contiguous segment
  int nseg;
  int chunks_per_seg;
  long long timeout_ns;
};

constexpr int kXchgThreads = 512;
constexpr int kXchgUnroll = 4;
constexpr int kXchgChunk = kXchgThreads * kXchgUnroll;  // vectors per work unit (32 KB)

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint4 ld_peer(const uint4* p) {
  uint4 v;  // .cv: never serve peer data from a stale L1 line
  asm volatile("ld.global.cv.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <bool kBf16>
__global__ void __launch_bounds__(kXchgThreads, 2) xchg_pull_kernel(XchgArgs a) {
  using Cv = Cvt<kBf16>;
  __shared__ unsigned s_epoch;
  __shared__ const uint4* s_src[MIMO_MAX_PEERS];
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();  // the source buffer was written by the previous kernel(s) of this stream
  if (tid == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(a.ctl);
  if (tid < a.G) s_src[tid] = static_cast<const uint4*>(a.src[tid]);
  __syncthreads();
  const unsigned e = s_epoch;
  if (tid < a.G) {
    if (blockIdx.x < 2) {  // announce (idempotent; two blocks so that it never depends on one block's scheduling)
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.ready[tid] + a.r), "r"(e) : "memory");
    }
    const unsigned* mine = a.ready[a.r] + tid;
    const unsigned long long t0 = globaltimer_ns();
    unsigned v;
    while (true) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (static_cast<int>(v - e) >= 0) break;
      __nanosleep(100);
      if (globaltimer_ns() - t0 > static_cast<unsigned long long>(a.timeout_ns)) {
        printf("mimo: exchange %u timed out waiting for peer %d (rank %d of %d saw %u)\n", e, tid, a.r, a.G, v);
        __trap();
      }
    }
  }
  __syncthreads();

  const int hwp = a.hw / a.G;
  const int F = a.fl * a.G;
  const int cv = a.C >> 3;  // vectors per row
  const int units = a.nseg * a.chunks_per_seg;
#pragma unroll 1
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int seg = u / a.chunks_per_seg;
    const int chunk = u - seg * a.chunks_per_seg;
    const int s = seg % a.G;
    const int k = (seg / a.G) % a.fl;
    const int bb = seg / (a.G * a.fl);
    long long src_row, dst_row;
    if (a.mode == 0) {  // frames -> pixels: pixel shard r of peer s's frame k  ->  my frame s * fl + k
      src_row = static_cast<long long>(bb * a.fl + k) * a.hw + static_cast<long long>(a.r) * hwp;
      dst_row = static_cast<long long>(bb * F + s * a.fl + k) * hwp;
    } else if (a.mode == 1) {  // pixels -> frames: my frame r * fl + k in peer s's pixel shard  ->  pixel shard s of frame k
      src_row = static_cast<long long>(bb * F + a.r * a.fl + k) * hwp;
      dst_row = static_cast<long long>(bb * a.fl + k) * a.hw + static_cast<long long>(s) * hwp;
    } else {  // all-gather: peer s's whole buffer -> my slot s
      src_row = 0;
      dst_row = 0;
    }
    const long long v0 = static_cast<long long>(chunk) * kXchgChunk;
    const uint4* sp = s_src[s] + src_row * cv + v0;
    long long doff = dst_row * cv + v0;
    if (a.mode == 2) doff += static_cast<long long>(s) * a.seg_vecs;
    uint4* dp = static_cast<uint4*>(a.dst) + doff;
    const uint4* rp = a.residual ? static_cast<const uint4*>(a.residual) + doff : nullptr;
    const int left = a.seg_vecs - static_cast<int>(v0);  // vectors of this segment from v0 on
    uint4 val[kXchgUnroll], res[kXchgUnroll];
#pragma unroll
    for (int i = 0; i < kXchgUnroll; ++i) {
      const int idx = tid + i * kXchgThreads;
      if (idx < left) val[i] = ld_peer(sp + idx);
    }
    if (rp) {
#pragma unroll
      for (int i = 0; i < kXchgUnroll; ++i) {
        const int idx = tid + i * kXchgThreads;
        if (idx < left) res[i] = rp[idx];
      }
    }
#pragma unroll
    for (int i = 0; i < kXchgUnroll; ++i) {
      const int idx = tid + i * kXchgThreads;
      if (idx < left) {
        uint4 o = val[i];
        if (rp) {  // out = residual + pulled, one rounding (the GEMM epilogue this replaces adds in fp32 as well)
          const uint32_t wv[4] = {val[i].x, val[i].y, val[i].z, val[i].w};
          const uint32_t wr[4] = {res[i].x, res[i].y, res[i].z, res[i].w};
          uint32_t wo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x = Cv::unpack(wv[j]);
            const float2 y = Cv::unpack(wr[j]);
            wo[j] = Cv::pack(x.x + y.x, x.y + y.y);
          }
          o = make_uint4(wo[0], wo[1], wo[2], wo[3]);
        }
        dp[idx] = o;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned t = atomicAdd(a.ctl + 1, 1u);
    if (t == gridDim.x - 1) {
      a.ctl[1] = 0;
      __threadfence();
      atomicExch(a.ctl, e + 1);
    }
  }
}

}  // namespace mimo

using namespace mimo;

extern "C" int mimo_exchange(const mimo_exchange_params* p, void* stream) {
  if (!p || !p->dst || !p->ctl) return set_error(MIMO_ERR_ARG, "mimo_exchange: null pointer");
  if (p->G < 1 || p->G > MIMO_MAX_PEERS || p->r < 0 || p->r >= p->G)
    return set_error(MIMO_ERR_ARG, "mimo_exchange: bad group size / rank");
  if (p->mode < 0 || p->mode > 2) return set_error(MIMO_ERR_ARG, "mimo_exchange: mode must be 0, 1 or 2");
  if (p->b <= 0 || p->fl <= 0 || p->hw <= 0 || p->C <= 0 || (p->C % 8))
    return set_error(MIMO_ERR_ARG, "mimo_exchange: bad sizes (C must be a multiple of 8)");
  if (p->mode != 2 && (p->hw % p->G)) return set_error(MIMO_ERR_ARG, "mimo_exchange: hw must be divisible by the group size");
  for (int s = 0; s < p->G; ++s)
    if (!p->peer_src[s] || !p->peer_ready[s]) return set_error(MIMO_ERR_ARG, "mimo_exchange: null peer pointer");
  if (int rc = ensure_device()) return rc;
  XchgArgs a;
  for (int s = 0; s < MIMO_MAX_PEERS; ++s) {
    a.src[s] = s < p->G ? p->peer_src[s] : nullptr;
    a.ready[s] = s < p->G ? static_cast<unsigned*>(p->peer_ready[s]) : nullptr;
  }
  a.ctl = static_cast<unsigned*>(p->ctl);
  a.dst = p->dst;
  a.residual = p->residual;
  a.mode = p->mode;
  a.G = p->G;
  a.r = p->r;
  a.b = p->b;
  a.fl = p->fl;
  a.hw = p->hw;
  a.C = p->C;
  long long seg_vecs, nseg;
  if (p->mode == 2) {  // b * fl * hw rows of C per rank, gathered rank-major
    seg_vecs = static_cast<long long>(p->b) * p->fl * p->hw * (p->C / 8);
    nseg = p->G;
    a.b = 1;
    a.fl = 1;
  } else {
    seg_vecs = static_cast<long long>(p->hw / p->G) * (p->C / 8);
    nseg = static_cast<long long>(p->b) * p->fl * p->G;
  }
  if (seg_vecs > 0x7fffffffLL || nseg > 0x7fffffffLL) return set_error(MIMO_ERR_ARG, "mimo_exchange: too large");
  a.seg_vecs = static_cast<int>(seg_vecs);
  a.nseg = static_cast<int>(nseg);
  a.chunks_per_seg = static_cast<int>((seg_vecs + kXchgChunk - 1) / kXchgChunk);
  a.timeout_ns = (p->timeout_ms > 0 ? p->timeout_ms : 30000) * 1000000LL;
  const long long units = nseg * a.chunks_per_seg;
  long long grid = p->max_blocks > 0 ? p->max_blocks : 2LL * num_sms();
  if (grid > units) grid = units;
  if (grid < 2) grid = 2;  // the two announcing blocks
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = p->dtype == MIMO_BF16
                      ? launch_k(xchg_pull_kernel<true>, dim3(static_cast<unsigned>(grid)), dim3(kXchgThreads), 0, st, a)
                      : launch_k(xchg_pull_kernel<false>, dim3(static_cast<unsigned>(grid)), dim3(kXchgThreads), 0, st, a);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error("exchange launch", e);
  return MIMO_OK;
}
