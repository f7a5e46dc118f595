// Host-side helpers shared by the C-ABI translation units: error reporting, device checks, TMA tensor-map
// encoding through the driver entry point (no link-time dependency on libcuda, so the library loads — and
// exports its symbols — on a machine without a GPU driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mimo {

int set_error(int code, const char* msg);
int set_cuda_error(const char* what, cudaError_t e);
// MIMO_OK if the current device is sm_100 (cached per process); error otherwise.
int ensure_device();
int num_sms();
// rank-`rank` tiled tensor map, 16-bit elements, 128-byte swizzle, zero OOB fill.
// dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride of dims[i+1].
int encode_tmap(CUtensorMap* out, int dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128);

inline unsigned div_up(long long a, long long b) { return static_cast<unsigned>((a + b - 1) / b); }

}  // namespace mimo
