// Bottleneck layers of wav2lip256 whose output map is 1x1 (face_encoder_blocks.7: 4x4 valid conv + 1x1 conv, face_decoder_blocks.0/1:
// 1x1 conv and the k4 ConvT on the 1x1 map, audio_encoder.11/12; avatars/wav2lip/models/wav2lip_v2.py:36-39, 56-58, 60-63):
// M = batch <= 16 rows, K = 512 .. 8192, Cout = 512 .. 8192 — 8 to 17 MB of weights per layer for 16 rows of work.  On the
// tensor-core kernels these were 12-22 us each (an M=128 tile holds 16 useful rows; split-K + finalize to find parallelism).
// They are pure weight streaming: one warp per output channel reads its K-major weight row once with 128-bit loads, the
// <= 16 input rows are staged through shared memory in 1024-element chunks, fp32 accumulation, bias + ReLU, fp16 out.
// HBM-bound: bytes = Cout * K * 2.
#include "ltb_internal.h"
#include "ptx_sm100.cuh"

namespace ltb {

constexpr int kFcRows = 16;
constexpr int kFcChunk = 1024;   // K elements staged per step: 16 rows x 1024 x 2 B = 32 KB

__global__ void __launch_bounds__(256) fc_rows_kernel(const FcParams p) {
  __shared__ __align__(16) __half xs[kFcRows][kFcChunk];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co = blockIdx.x * 8 + warp;
  pdl_launch_dependents();
  pdl_wait();
  float acc[kFcRows];
#pragma unroll
  for (int m = 0; m < kFcRows; ++m) acc[m] = 0.f;
  const __half* wrow = p.w + (size_t)(co < p.Cout ? co : 0) * p.K;
  for (int k0 = 0; k0 < p.K; k0 += kFcChunk) {
    const int kc = min(kFcChunk, p.K - k0);
    __syncthreads();   // previous chunk fully consumed
    // stage rows [0, M) x [k0, k0 + kc): K index = tap * Cin + c  ->  input element tap * tap_pitch + c of the row
    for (int v = threadIdx.x; v < kFcRows * (kFcChunk / 8); v += 256) {
      const int m = v / (kFcChunk / 8), kk = (v % (kFcChunk / 8)) * 8;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (m < p.M && kk < kc) {
        const int k = k0 + kk;
        const int tap = k / p.Cin, c = k - tap * p.Cin;
        val = *reinterpret_cast<const uint4*>(p.in + (size_t)m * p.in_row + (size_t)tap * p.tap_pitch + c);
      }
      *reinterpret_cast<uint4*>(&xs[m][kk]) = val;
    }
    __syncthreads();
    if (co < p.Cout) {
      for (int kk = lane * 8; kk < kc; kk += 256) {
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wrow + k0 + kk));
        const __half2* wh = reinterpret_cast<const __half2*>(&wv);
        const float2 w0 = __half22float2(wh[0]), w1 = __half22float2(wh[1]), w2 = __half22float2(wh[2]), w3 = __half22float2(wh[3]);
#pragma unroll
        for (int m = 0; m < kFcRows; ++m) {
          const uint4 xv = *reinterpret_cast<const uint4*>(&xs[m][kk]);
          const __half2* xh = reinterpret_cast<const __half2*>(&xv);
          const float2 x0 = __half22float2(xh[0]), x1 = __half22float2(xh[1]), x2 = __half22float2(xh[2]), x3 = __half22float2(xh[3]);
          float a = acc[m];
          a = fmaf(x0.x, w0.x, a);
          a = fmaf(x0.y, w0.y, a);
          a = fmaf(x1.x, w1.x, a);
          a = fmaf(x1.y, w1.y, a);
          a = fmaf(x2.x, w2.x, a);
          a = fmaf(x2.y, w2.y, a);
          a = fmaf(x3.x, w3.x, a);
          a = fmaf(x3.y, w3.y, a);
          acc[m] = a;
        }
      }
    }
  }
  if (co >= p.Cout) return;
  // warp reduction: after the butterfly every lane holds every row's total; lane m writes row m
  float mine = 0.f;
#pragma unroll
  for (int m = 0; m < kFcRows; ++m) {
    float a = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == m) mine = a;
  }
  if (lane < p.M) {
    float v = mine + __ldg(p.bias + co);
    if (p.relu) v = fmaxf(v, 0.f);
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    p.out[(size_t)lane * p.out_row + co] = __float2half_rn(v);
  }
}

// a conv whose output map is 1x1 and whose taps tile the whole (unpadded) input map, batch <= 16
bool fc_rows_supported(const ConvParams& c) {
  if (c.nphases != 1 || c.GH != 1 || c.GW != 1 || c.OH != 1 || c.OW != 1 || c.N > kFcRows || c.res || c.zbatch > 1) return false;
  if (c.Cin % 8 || c.ICtot % 8 || c.ic_off % 8 || c.Ktot != c.ph[0].ntaps * c.Cin || c.ph[0].ntaps != c.IH * c.IW) return false;
  for (int t = 0; t < c.ph[0].ntaps; ++t)
    if (c.ph[0].dy[t] != t / c.IW || c.ph[0].dx[t] != t % c.IW) return false;
  return c.Cout >= 64;
}

void fc_rows_make(const ConvParams& c, FcParams* f) {
  f->in = c.in + c.ic_off;
  f->w = c.w + c.ph[0].koff;
  f->bias = c.bias;
  f->out = c.out + c.oc_off;
  f->M = c.N;
  f->K = c.Ktot;
  f->Cin = c.Cin;
  f->Cout = c.Cout;
  f->relu = c.relu;
  f->tap_pitch = c.ICtot;
  f->in_row = (long long)c.IH * c.IW * c.ICtot;
  f->out_row = c.OCtot;
}

cudaError_t launch_fc_rows(const FcParams& f, cudaStream_t st) {
  return launch_kernel_pdl(fc_rows_kernel, dim3((f.Cout + 7) / 8), dim3(256), 0, st, f);
}

}  // namespace ltb
