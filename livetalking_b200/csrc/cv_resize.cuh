// OpenCV 8-bit INTER_LINEAR resize arithmetic shared by the paste-back kernels (paste.cu: wav2lip, ultralight.cu: UltraLight):
// bit-exact with cv2.resize — 11-bit fixed-point taps computed with the same float32/float64 expression order.
#pragma once
#include <cstdint>

namespace ltb {

__device__ __forceinline__ int mirror_index_p(int size, int index) {
  const int turn = index / size, res = index % size;
  return (turn % 2 == 0) ? res : size - res - 1;
}

// OpenCV linear-resize tap for destination index d: source index s (clamped) and fixed-point weights
__device__ __forceinline__ void cv_tap(int d, double scale, int src_len, bool clamp_taps, int& s, int& w0, int& w1) {
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  int si = (int)floorf(f);
  f = __fsub_rn(f, (float)si);
  if (clamp_taps) {  // horizontal: resize.cpp clamps the tap position and zeroes the fraction
    if (si < 0) {
      si = 0;
      f = 0.f;
    }
    if (si >= src_len - 1) {
      si = src_len - 1;
      f = 0.f;
    }
  }
  s = si;
  w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  w1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

__device__ __forceinline__ int trunc_u8(float v) { return (int)(unsigned char)v; }  // ndarray.astype(np.uint8)

}  // namespace ltb
