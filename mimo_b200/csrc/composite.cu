// Scene compositing of run_edit.py (:282-300), the per-frame blend chain as ONE pass over the frame:
//   res = canvas * mask + bk * (1 - mask)                  (numpy: uint8 * float32 -> float32 arithmetic)
//   res = res * (1 - occ / 255.0) + vid * (occ / 255.0)    (optional; numpy promotes to float64 here)
//   out = prev * (1 - factor) + res * factor               (optional cross-fade of overlapping clips; float64)
//   out.astype(uint8)                                      (truncation toward zero)
// Every intermediate is computed in the type numpy computes it in, so the bytes match the reference's exactly.
// HBM-bound: 3 uint8 images + 1 float mask in, 1 uint8 image out; one thread per pixel (3 channels).
#include <cuda_runtime.h>

#include "../../include/mimo_b200.h"
#include "host_util.h"

namespace mimo {

__global__ void composite_kernel(const uint8_t* __restrict__ canvas, const uint8_t* __restrict__ bk,
                                 const float* __restrict__ mask, const uint8_t* __restrict__ occ,
                                 const uint8_t* __restrict__ vid, const uint8_t* __restrict__ prev, double factor,
                                 uint8_t* __restrict__ out, long long pixels) {
  const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (p >= pixels) return;
  const float m = mask[p];
  const float om = 1.0f - m;
  const float ff = static_cast<float>(factor);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const long long i = p * 3 + c;
    const float r32 = __fadd_rn(__fmul_rn(static_cast<float>(canvas[i]), m), __fmul_rn(static_cast<float>(bk[i]), om));
    double r;
    if (occ) {
      const double o = static_cast<double>(occ[p]) / 255.0;
      r = __dadd_rn(__dmul_rn(static_cast<double>(r32), 1.0 - o), __dmul_rn(static_cast<double>(vid[i]), o));
      if (prev) r = __dadd_rn(__dmul_rn(static_cast<double>(prev[i]), 1.0 - factor), __dmul_rn(r, factor));
    } else if (prev) {
      // res is still float32 here: numpy multiplies a float32 array by the Python float in float32
      r = __dadd_rn(__dmul_rn(static_cast<double>(prev[i]), 1.0 - factor), static_cast<double>(__fmul_rn(r32, ff)));
    } else {
      r = static_cast<double>(r32);
    }
    out[i] = static_cast<uint8_t>(static_cast<int>(r));  // astype(np.uint8): truncation
  }
}

}  // namespace mimo

using namespace mimo;

extern "C" int mimo_composite_frame(const void* canvas, const void* bk, const float* mask, const void* occ,
                                    const void* vid, const void* prev, double factor, void* out, int64_t pixels,
                                    void* stream) {
  if (!canvas || !bk || !mask || !out || pixels <= 0) return set_error(MIMO_ERR_ARG, "mimo_composite_frame: bad arguments");
  if ((occ == nullptr) != (vid == nullptr))
    return set_error(MIMO_ERR_ARG, "mimo_composite_frame: occlusion mask and original frame come together");
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = div_up(pixels, 256);
  composite_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint8_t*>(canvas), static_cast<const uint8_t*>(bk), mask,
                                           static_cast<const uint8_t*>(occ), static_cast<const uint8_t*>(vid),
                                           static_cast<const uint8_t*>(prev), factor, static_cast<uint8_t*>(out), pixels);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error("composite launch", e);
  return MIMO_OK;
}
