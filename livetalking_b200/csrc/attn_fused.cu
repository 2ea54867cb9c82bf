// Fused multi-head attention on tcgen05 (sm_100a): softmax(scale * Q K^T) V with the score matrix S living only in TMEM and
// shared memory — replaces the three-kernel path (batched QK^T GEMM -> fp16 S in HBM -> softmax kernel -> batched PV GEMM)
// of the diffusers Attention blocks (UNet self / cross attention, avatars/musetalk/models/unet.py:29-48 -> diffusers
// BasicTransformerBlock) and the Whisper encoder layers (avatars/musetalk/whisper/audio2feature.py:106-117).
//
// One CTA = 128 queries of one (batch, head).  TWO passes over the keys instead of an online rescale of O:
//   pass 1: S_t = Q K_t^T (tcgen05.mma, M=128 queries, N=128 keys, K = head dim) -> TMEM; the softmax warps read their row
//           (one query per thread, one TMEM lane) and keep the running maximum m and the sum l = sum exp(scale*(s - m));
//   pass 2: S_t is recomputed (the head dim is 48..160: re-running 3..10 K steps is cheaper than keeping or rescaling
//           anything), P_t = exp(scale*(s - m)) / l is written as fp16 into shared memory in the canonical K-major 128B-swizzled
//           layout and consumed as the A operand of O += P_t V_t (M=128, N = head dim, K=128 keys) — O accumulates in TMEM.
// Q / K tiles come by TMA straight out of the fused qkv (or q / kv) projection buffers (4-D maps (d, token, head, batch); the
// box is 64 channels wide and the tensor's channel extent is the head dim, so the padding up to 64 is zero-filled); V^T tiles from
// the [B*H][d][keys] buffer transpose_heads writes.  TMEM: 128 columns of S + d columns of O (<= 256 for d <= 128: two CTAs per SM
// overlap each other's MMA / softmax phases, so the per-CTA pipeline is kept simple and serial).
// Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 softmax + epilogue: TMEM lane quarter = warp % 4, and the two
// warps of a quarter split the 128 key columns of a tile (ncu r02q: with four softmax warps the kernel sat at 32 % of the SFU pipe and
// 37 % issue utilisation — latency bound on one warp per scheduler); their row statistics are merged once, after pass 1.
#include <cuda.h>

#include <mutex>

#include "ltb_internal.h"
#include "ops.h"
#include "ptx_sm100.cuh"

namespace ltb {

struct alignas(64) AttnParams {
  CUtensorMap tm_q;    // 4-D (d, nq, H, B), box (64, 128, 1, 1), SWIZZLE_128B
  CUtensorMap tm_k;    // 4-D (d, nk_rows, H, B), box (64, 128, 1, 1)
  CUtensorMap tm_vt;   // 3-D (n_pad, d, B*H), box (64, d, 1)
  __half* out;         // [B*nq][out_pitch], head h at columns [h*d, (h+1)*d)
  int out_pitch, nq, valid, d, H, kstages;
  float scale_log2e;   // scale * log2(e): probabilities are exp2((s - m) * scale_log2e)
};

constexpr int kAttnThreads = 320;
constexpr int kTileBytes = 128 * 128;   // 128 rows x 64 fp16

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads) attn_fused_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full, v_empty, s_full, s_empty, p_full, p_empty, o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 row_stat[2][128];      // (max, sum) of each query row over one half of the key columns

  pdl_launch_dependents();   // the output projection (a PDL-launched conv kernel) may start its prologue under this kernel's tail
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 128, bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
  const int chunks = (p.d + 63) / 64;                 // 64-channel K chunks of the QK^T contraction
  const int nkt = (p.valid + 127) / 128;              // key tiles (keys >= valid are masked; rows beyond the tensor come back as zeros)
  const uint32_t KS = (uint32_t)p.kstages;            // K ring depth (1 when the head dim needs three chunks: shared memory)
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem0;                                       // chunks x [128 x 128 B]
  const uint32_t k_smem = q_smem + chunks * kTileBytes;                // 2 stages x chunks x [128 x 128 B]
  const uint32_t v_smem = k_smem + KS * chunks * kTileBytes;           // 2 key blocks x [d rows x 128 B]
  const uint32_t v_blk = ((uint32_t)p.d * 128u + 1023u) & ~1023u;
  const uint32_t p_smem = v_smem + 2 * v_blk;                          // 2 key blocks x [128 x 128 B]
  uint8_t* const p_ptr = smem_raw + (p_smem - smem_u32(smem_raw));
  const uint32_t tcols = (128 + p.d <= 256) ? 256u : 512u;

  if (tid == 0) {
    mbar_init(smem_u32(&q_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&k_full[s]), 1);
      mbar_init(smem_u32(&k_empty[s]), 1);
    }
    mbar_init(smem_u32(&v_full), 1);
    mbar_init(smem_u32(&v_empty), 1);
    mbar_init(smem_u32(&s_full), 1);
    mbar_init(smem_u32(&s_empty), 8);     // the eight softmax warps
    mbar_init(smem_u32(&p_full), 8);
    mbar_init(smem_u32(&p_empty), 1);
    mbar_init(smem_u32(&o_full), 1);
    mbar_fence_init();
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_vt);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), tcols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_s = tmem, tmem_o = tmem + 128;

  if (warp == 0) {
    // =============================================================== TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(&q_full), chunks * kTileBytes);
      for (int c = 0; c < chunks; ++c) tma_load_4d(q_smem + c * kTileBytes, &p.tm_q, smem_u32(&q_full), c * 64, q0, h, b);
      uint32_t ki = 0;
      for (int pass = 0; pass < 2; ++pass)
        for (int kt = 0; kt < nkt; ++kt, ++ki) {
          const uint32_t s = ki % KS;
          mbar_wait(smem_u32(&k_empty[s]), ((ki / KS) & 1u) ^ 1u);
          mbar_arrive_expect_tx(smem_u32(&k_full[s]), chunks * kTileBytes);
          for (int c = 0; c < chunks; ++c)
            tma_load_4d(k_smem + (s * chunks + c) * kTileBytes, &p.tm_k, smem_u32(&k_full[s]), c * 64, kt * 128, h, b);
          if (pass == 1) {
            mbar_wait(smem_u32(&v_empty), (kt & 1u) ^ 1u);
            mbar_arrive_expect_tx(smem_u32(&v_full), 2 * p.d * 128);
            for (int j = 0; j < 2; ++j) tma_load_3d(v_smem + j * v_blk, &p.tm_vt, smem_u32(&v_full), kt * 128 + j * 64, 0, bh);
          }
        }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer (warp-uniform, elected-lane predication)
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t idesc_s = umma_idesc_f16(128, 128);
    const uint32_t idesc_o = umma_idesc_f16(128, p.d);
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
    const int ksteps_last = ((p.d - 1) % 64) / 16 + 1;
    mbar_wait(smem_u32(&q_full), 0);
    uint32_t ki = 0, si = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int kt = 0; kt < nkt; ++kt, ++ki, ++si) {
        const uint32_t s = ki % KS;
        mbar_wait(smem_u32(&k_full[s]), (ki / KS) & 1u);
        mbar_wait(smem_u32(&s_empty), (si & 1u) ^ 1u);      // the softmax warps have read the previous S tile
        tc_fence_after();
        for (int c = 0; c < chunks; ++c) {
          const uint32_t a_lo = (((q_smem + c * kTileBytes) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t b_lo = (((k_smem + (s * chunks + c) * kTileBytes) & 0x3FFFFu) >> 4) | (1u << 16);
          const int ks = (c == chunks - 1) ? ksteps_last : 4;
          for (int k = 0; k < ks; ++k)
            umma_f16_lohi_if(leader, tmem_s, a_lo + k * 2, kDescHi, b_lo + k * 2, kDescHi, idesc_s, (c | k) ? 1u : 0u);
        }
        umma_commit_if(leader, smem_u32(&k_empty[s]));
        umma_commit_if(leader, smem_u32(&s_full));
        if (pass == 1) {
          // O += P_kt V_kt : A = the fp16 probabilities the softmax warps wrote, B = V^T tile (d rows x 128 keys)
          mbar_wait(smem_u32(&p_full), kt & 1u);
          mbar_wait(smem_u32(&v_full), kt & 1u);
          tc_fence_after();
          for (int j = 0; j < 2; ++j) {
            const uint32_t a_lo = (((p_smem + j * kTileBytes) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo = (((v_smem + j * v_blk) & 0x3FFFFu) >> 4) | (1u << 16);
            for (int k = 0; k < 4; ++k)
              umma_f16_lohi_if(leader, tmem_o, a_lo + k * 2, kDescHi, b_lo + k * 2, kDescHi, idesc_o, (kt | j | k) ? 1u : 0u);
          }
          umma_commit_if(leader, smem_u32(&p_empty));
          umma_commit_if(leader, smem_u32(&v_empty));
        }
      }
    umma_commit_if(leader, smem_u32(&o_full));
    __syncwarp();
  } else {
    // =============================================================== softmax + epilogue: one query row x 64 key columns per thread
    const int q = warp & 3;                        // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;              // which 64 key columns of every tile
    const int row = q * 32 + lane;                 // query row inside the tile
    const uint32_t trow = ((uint32_t)(q * 32) << 16);
    const int cbase = half * 64;
    float m = -INFINITY, l = 0.f;
    uint32_t si = 0;
    // ---- pass 1: running maximum and sum over this thread's columns
    for (int kt = 0; kt < nkt; ++kt, ++si) {
      mbar_wait(smem_u32(&s_full), si & 1u);
      tc_fence_after();
      const int lim0 = p.valid - kt * 128 - cbase;     // columns [0, lim0) of this thread's 64 are real keys
      if (lim0 > 0) {
        uint32_t v0[32], v1[32];
        tmem_ld32(tmem_s + trow + cbase, v0);
        tmem_ld32(tmem_s + trow + cbase + 32, v1);
        tmem_ld_wait();
        float tmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < lim0) tmax = fmaxf(tmax, __uint_as_float(v0[i]));
          if (i + 32 < lim0) tmax = fmaxf(tmax, __uint_as_float(v1[i]));
        }
        const float m_new = fmaxf(m, tmax);
        float lsum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < lim0) lsum += ex2f((__uint_as_float(v0[i]) - m_new) * p.scale_log2e);
          if (i + 32 < lim0) lsum += ex2f((__uint_as_float(v1[i]) - m_new) * p.scale_log2e);
        }
        l = l * ex2f((m - m_new) * p.scale_log2e) + lsum;   // m = -inf on the first tile: ex2(-inf) = 0
        m = m_new;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty));
    }
    // ---- merge the two column halves of every row (named barrier over the 256 softmax threads)
    row_stat[half][row] = make_float2(m, l);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    {
      const float2 o = row_stat[half ^ 1][row];
      const float mm = fmaxf(m, o.x);
      // a half with no valid key at all has m = -inf, l = 0: its term vanishes (ex2(-inf) = 0; mm is finite since valid >= 1)
      l = (m == -INFINITY ? 0.f : l * ex2f((m - mm) * p.scale_log2e)) + (o.x == -INFINITY ? 0.f : o.y * ex2f((o.x - mm) * p.scale_log2e));
      m = mm;
    }
    const float inv_l = 1.f / l;
    // ---- pass 2: probabilities -> shared memory (A operand of the PV product); this thread fills rows of key block `half`
    for (int kt = 0; kt < nkt; ++kt, ++si) {
      mbar_wait(smem_u32(&s_full), si & 1u);
      tc_fence_after();
      mbar_wait(smem_u32(&p_empty), (kt & 1u) ^ 1u);     // the PV MMAs of the previous tile have consumed P
      const int lim0 = p.valid - kt * 128 - cbase;
      uint8_t* prow = p_ptr + half * kTileBytes + row * 128;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        if (lim0 > c0) {
          tmem_ld32(tmem_s + trow + cbase + c0, v);
          tmem_ld_wait();
        }
        const int lim = lim0 - c0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {          // 8 keys = one 16-byte chunk of the K-major row
          uint32_t w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = g * 8 + 2 * u;
            const float a = (i < lim) ? ex2f((__uint_as_float(v[i]) - m) * p.scale_log2e) * inv_l : 0.f;
            const float c = (i + 1 < lim) ? ex2f((__uint_as_float(v[i + 1]) - m) * p.scale_log2e) * inv_l : 0.f;
            w[u] = f32x2_to_f16x2_sat(a, c);
          }
          const int chunk = (c0 >> 3) + g;                  // 0..7 inside the 64-key block
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();        // generic-proxy writes of P visible to the tensor core's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&p_full));
        mbar_arrive(smem_u32(&s_empty));
      }
    }
    // ---- epilogue: O (fp32, TMEM) -> fp16 rows of the output
    mbar_wait(smem_u32(&o_full), 0);
    tc_fence_after();
    const int qi = q0 + row;
    __half* orow = p.out + ((size_t)b * p.nq + qi) * p.out_pitch + h * p.d;
    for (int c0 = half * 16; c0 < p.d; c0 += 32) {     // the two warps of a quarter alternate 16-column groups
      uint32_t v[16];
      tmem_ld16(tmem_o + trow + c0, v);
      tmem_ld_wait();
      if (qi < p.nq) {
        uint4 o0, o1;
        o0.x = f32x2_to_f16x2_sat(__uint_as_float(v[0]), __uint_as_float(v[1]));
        o0.y = f32x2_to_f16x2_sat(__uint_as_float(v[2]), __uint_as_float(v[3]));
        o0.z = f32x2_to_f16x2_sat(__uint_as_float(v[4]), __uint_as_float(v[5]));
        o0.w = f32x2_to_f16x2_sat(__uint_as_float(v[6]), __uint_as_float(v[7]));
        o1.x = f32x2_to_f16x2_sat(__uint_as_float(v[8]), __uint_as_float(v[9]));
        o1.y = f32x2_to_f16x2_sat(__uint_as_float(v[10]), __uint_as_float(v[11]));
        o1.z = f32x2_to_f16x2_sat(__uint_as_float(v[12]), __uint_as_float(v[13]));
        o1.w = f32x2_to_f16x2_sat(__uint_as_float(v[14]), __uint_as_float(v[15]));
        *reinterpret_cast<uint4*>(orow + c0) = o0;
        *reinterpret_cast<uint4*>(orow + c0 + 8) = o1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, tcols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 attn_encode_fn() {
  static EncodeTiledFn2 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(ptr);
  });
  return fn;
}

static bool attn_encode(CUtensorMap* tm, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
  EncodeTiledFn2 fn = attn_encode_fn();
  if (!fn) return false;
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool attn_fused_supported(int d, int q_pitch, int kv_pitch, int n_pad) {
  return d >= 16 && d <= 160 && d % 16 == 0 && q_pitch % 8 == 0 && kv_pitch % 8 == 0 && n_pad % 8 == 0 && attn_encode_fn() != nullptr;
}

// q: [B][nq] rows of q_pitch elements, head h at columns [h*d, (h+1)*d); k: [B][kv_rows] rows of kv_pitch elements, same head layout;
// vt: [B*H][d][n_pad] (transpose_heads: zero for key >= kv_rows); keys [valid, ..) get probability 0.
cudaError_t launch_attn_fused(const __half* q, int q_pitch, const __half* k, int kv_pitch, int kv_rows, const __half* vt, int n_pad, int B, int H,
                              int nq, int valid, int d, float scale, __half* out, int out_pitch, cudaStream_t st) {
  if (!attn_fused_supported(d, q_pitch, kv_pitch, n_pad) || nq < 1 || valid < 1 || valid > kv_rows || valid > n_pad || out_pitch % 8)
    return cudaErrorInvalidValue;
  const int nk = kv_rows;
  AttnParams p;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)nq, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)q_pitch * 2, (cuuint64_t)d * 2, (cuuint64_t)nq * q_pitch * 2};
    cuuint32_t box[4] = {64, 128, 1, 1};
    if (!attn_encode(&p.tm_q, 4, q, dims, strides, box)) return cudaErrorInvalidValue;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)nk, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)kv_pitch * 2, (cuuint64_t)d * 2, (cuuint64_t)nk * kv_pitch * 2};
    cuuint32_t box[4] = {64, 128, 1, 1};
    if (!attn_encode(&p.tm_k, 4, k, dims, strides, box)) return cudaErrorInvalidValue;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)n_pad, (cuuint64_t)d, (cuuint64_t)B * H};
    cuuint64_t strides[2] = {(cuuint64_t)n_pad * 2, (cuuint64_t)d * n_pad * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)d, 1};
    if (!attn_encode(&p.tm_vt, 3, vt, dims, strides, box)) return cudaErrorInvalidValue;
  }
  p.out = out;
  p.out_pitch = out_pitch;
  p.nq = nq;
  p.valid = valid;
  p.d = d;
  p.H = H;
  p.scale_log2e = scale * 1.4426950408889634f;
  const int chunks = (d + 63) / 64;
  const int v_blk = (d * 128 + 1023) & ~1023;
  p.kstages = chunks <= 2 ? 2 : 1;
  const int smem = (1 + p.kstages) * chunks * kTileBytes + 2 * v_blk + 2 * kTileBytes + 1024;
  constexpr int kMaxSmem = 200 * 1024;
  static SmemConfigOnce once;
  if (cudaError_t e = once.ensure(attn_fused_kernel, kMaxSmem); e != cudaSuccess) return e;
  if (smem > kMaxSmem) return cudaErrorInvalidValue;
  attn_fused_kernel<<<dim3((nq + 127) / 128, B * H), kAttnThreads, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace ltb
