// Tensor-core stem of wav2lip256 (stem_umma.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ltb {

struct alignas(64) StemParams {
  CUtensorMap tm_in;  // 4-D (8 ch, 264, 262, B) zero-bordered image, box (8, 16, 22, 1), SWIZZLE_NONE
  CUtensorMap tm_w;   // 3-D (64 k, 16 cout, 7 kernel rows), box = whole tensor, SWIZZLE_128B
  __half* out;
  const float* bias;
  int OCtot, oc_off;
  int total_tiles;
};

int stem_make_plan(const __half* img_pad, int B, const __half* w_tap_major, const float* bias, __half* out, int OCtot, int oc_off,
                   StemParams* sp);
cudaError_t launch_stem(const StemParams& sp, cudaStream_t st);

}  // namespace ltb
