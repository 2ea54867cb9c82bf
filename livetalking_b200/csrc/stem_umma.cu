// wav2lip256 stem on tensor cores: Conv2d(6,16,k7,s1,p3)+BN+ReLU (avatars/wav2lip/models/wav2lip_v2.py:13) over the
// zero-bordered 8-channel fp16 image written by w2l_prep_faces ([B,262,264,8]: 3 masked + 3 full + 2 zero channels).
//
// With 8 channels a pixel is exactly one 16-byte K chunk, so the im2col row of output pixel (y,x) for kernel row kh —
// 8 consecutive pixels x 8 channels = 64 K values (7 real taps + 1 zero-weighted) — is 128 CONTIGUOUS bytes of the image
// row, and the rows of neighbouring output pixels overlap by 112 bytes.  tcgen05's SWIZZLE_NONE K-major descriptor
// expresses that directly: core-matrix rows 16 B apart (implicit), K chunks LBO = 16 B apart, 8-row groups (one image row
// of the 16x8 output tile) SBO = 256 B apart (verified by ltb_umma_probe_noswz / tests/probe_umma.py).  So ONE
// 22 x 16 pixel halo (5.6 KB, plain TMA box) feeds all 7 x 4 MMAs of a 128-pixel tile: 20x less operand traffic than
// gathering seven 16 KB im2col tiles.  Persistent CTAs, weights (14 KB) resident, TMEM double-buffered.
#include <cuda.h>

#include <cstring>
#include <mutex>

#include "ltb_internal.h"
#include "ptx_sm100.cuh"
#include "stem_umma.h"

namespace ltb {

constexpr int kStemStages = 6;
constexpr int kStemABytes = 22 * 16 * 16;   // 5632 (halo: 22 rows x 16 px x 8 ch fp16)
constexpr int kStemAStride = 6144;          // stage pitch
constexpr int kStemWBytes = 7 * 16 * 128;   // 7 kernel rows x 16 cout x 64 k

__global__ void __launch_bounds__(192, 1) stem_umma_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kStemStages], a_empty[kStemStages];
  __shared__ __align__(8) uint64_t w_full, acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_smem = smem0;                               // 14 KB, SWIZZLE_128B tiles of 16 rows
  const uint32_t a_smem = smem0 + kStemWBytes;                 // stages, SWIZZLE_NONE
  if (tid == 0) {
    for (int s = 0; s < kStemStages; ++s) {
      mbar_init(smem_u32(&a_full[s]), 1);
      mbar_init(smem_u32(&a_empty[s]), 1);
    }
    mbar_init(smem_u32(&w_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&acc_full[s]), 1);
      mbar_init(smem_u32(&acc_empty[s]), 4);
    }
    mbar_fence_init();
    tma_prefetch_desc(&p.tm_in);
    tma_prefetch_desc(&p.tm_w);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int tiles_per_img = 16 * 32;  // 256/16 x 256/8

  pdl_launch_dependents();   // PDL: weights (constants) before the wait, the padded image after it
  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(smem_u32(&w_full), kStemWBytes);
    tma_load_3d(w_smem, &p.tm_w, smem_u32(&w_full), 0, 0, 0);
  }
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ai = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ai) {
        const int img = t / tiles_per_img, r = t - img * tiles_per_img;
        const int ty = r >> 5, tx = r & 31;
        const uint32_t as = ai % kStemStages;
        mbar_wait(smem_u32(&a_empty[as]), ((ai / kStemStages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(smem_u32(&a_full[as]), kStemABytes);
        // padded image coordinates: output (y,x) reads padded rows y..y+6, padded pixels x..x+7
        tma_load_4d(a_smem + as * kStemAStride, &p.tm_in, smem_u32(&a_full[as]), 0, tx * 8, ty * 16, img);
      }
    }
  } else if (warp == 1) {
    {  // warp-uniform MMA issue (see umma_f16_lohi_if)
      const uint32_t leader = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = umma_idesc_f16(128, 16);
      // A: SWIZZLE_NONE, LBO = 16 B (next K chunk = next pixel), SBO = 256 B (next image row of the halo)
      constexpr uint32_t a_hi = (256u >> 4) | (1u << 14);
      constexpr uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      mbar_wait(smem_u32(&w_full), 0);
      uint32_t ai = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ai) {
        const uint32_t buf = ai & 1u;
        mbar_wait(smem_u32(&acc_empty[buf]), ((ai >> 1) & 1u) ^ 1u);
        const uint32_t as = ai % kStemStages;
        mbar_wait(smem_u32(&a_full[as]), (ai / kStemStages) & 1u);
        tc_fence_after();
        const uint32_t a_base = a_smem + as * kStemAStride;
#pragma unroll
        for (int kh = 0; kh < 7; ++kh) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a_lo = (((a_base + kh * 256 + k * 32) & 0x3FFFFu) >> 4) | (1u << 16);   // LBO = 16 B
            const uint32_t b_lo = (((w_smem + kh * 2048 + k * 32) & 0x3FFFFu) >> 4) | (1u << 16);
            umma_f16_lohi_if(leader, tmem + buf * 16, a_lo, a_hi, b_lo, b_hi, idesc, (kh | k) ? 1u : 0u);
          }
        }
        umma_commit_if(leader, smem_u32(&a_empty[as]));
        umma_commit_if(leader, smem_u32(&acc_full[buf]));
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane, ry = row >> 3, rx = row & 7;
    uint32_t ai = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ai) {
      const int img = t / tiles_per_img, r = t - img * tiles_per_img;
      const int ty = r >> 5, tx = r & 31;
      const uint32_t buf = ai & 1u;
      mbar_wait(smem_u32(&acc_full[buf]), (ai >> 1) & 1u);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(tmem + buf * 16 + ((uint32_t)(q * 32) << 16), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
      const size_t pix = ((size_t)img * 256 + ty * 16 + ry) * 256 + tx * 8 + rx;
      __half* optr = p.out + pix * p.OCtot + p.oc_off;
      uint4 ov[2];
      __half2* oh = reinterpret_cast<__half2*>(ov);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = fmaxf(__uint_as_float(v[2 * j]) + __ldg(p.bias + 2 * j), 0.f);
        const float y = fmaxf(__uint_as_float(v[2 * j + 1]) + __ldg(p.bias + 2 * j + 1), 0.f);
        oh[j] = __floats2half2_rn(fminf(x, 65504.f), fminf(y, 65504.f));
      }
      reinterpret_cast<uint4*>(optr)[0] = ov[0];
      reinterpret_cast<uint4*>(optr)[1] = ov[1];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int stem_make_plan(const __half* img_pad, int B, const __half* w_tap_major, const float* bias, __half* out, int OCtot, int oc_off,
                   StemParams* sp) {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  });
  if (!fn) return 1;
  std::memset(sp, 0, sizeof(*sp));
  cuuint32_t es[4] = {1, 1, 1, 1};
  {
    cuuint64_t dims[4] = {8, 264, 262, (cuuint64_t)B};
    cuuint64_t strides[3] = {16, 264 * 16, (cuuint64_t)262 * 264 * 16};
    cuuint32_t box[4] = {8, 16, 22, 1};
    if (fn(&sp->tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(img_pad), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 2;
  }
  {
    cuuint64_t dims[3] = {64, 16, 7};
    cuuint64_t strides[2] = {128, 16 * 128};
    cuuint32_t box[3] = {64, 16, 7};
    if (fn(&sp->tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(w_tap_major), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 2;
  }
  sp->out = out;
  sp->bias = bias;
  sp->OCtot = OCtot;
  sp->oc_off = oc_off;
  sp->total_tiles = B * 16 * 32;
  return 0;
}

cudaError_t launch_stem(const StemParams& sp, cudaStream_t st) {
  static SmemConfigOnce once;
  static std::atomic<int> sms_cached{0};
  constexpr int smem = kStemWBytes + kStemStages * kStemAStride + 1024;
  if (cudaError_t e = once.ensure(stem_umma_kernel, smem); e != cudaSuccess) return e;
  int sms = sms_cached.load();
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    sms_cached.store(sms);
  }
  const int grid = sp.total_tiles < sms ? sp.total_tiles : sms;
  return launch_kernel_pdl(stem_umma_kernel, dim3(grid), dim3(192), smem, st, sp);
}

}  // namespace ltb
