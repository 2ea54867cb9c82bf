// "y-stacked" 3x3 convolution for NARROW layers (Cout = 32 or 64) on tcgen05 (sm_100a) — wav2lip256's 64-channel 256x256
// decoder convs, the 80->32 output conv with the fused 1x1 head, the 32/64-channel encoder convs
// (avatars/wav2lip/models/wav2lip_v2.py:15-22, 83-91; Conv2d+BN(+residual)+ReLU blocks of conv.py:5-19).
//
// Why: one M=128,K=16 tcgen05.mma costs ~55 + 0.2*N cycles — at N = 64 / 32 the 4 KB A-operand fetch from shared memory
// dominates, so the halo kernel's nine N=BN instructions per K step run the tensor pipe at 50 % / 28 % of its rate
// (profiles/r01n_diag_halo.txt: L52 is MMA-issue bound at 70 us against a 41 us HBM floor).  Here the three taps of one
// kernel COLUMN (dy = 0,1,2 at fixed dx) are stacked along N: one instruction per (dx, K step) with N = 3*BN produces three
// partial sums
//     D_dy[y, x] = sum_dx sum_k in[y, x+dx-1, k] * W[dy, dx][k, :]          (A view = halo shifted by dx only)
// and the epilogue combines   out[y, x] = D_0[y-1, x] + D_1[y, x] + D_2[y+1, x].
// A row of the tile is 8 TMEM lanes, so y+-1 is lane +-8: a warp shuffle inside a 32-lane quarter, and a 2 KB shared-memory
// exchange between neighbouring quarters (warps can only read their own TMEM lane quarter).  The first and last MMA row of
// a tile have no neighbour: tiles overlap by two rows (16*NSUB - 2 output rows per tile), ~7-19 % extra MMA rows against
// a 3x cut in instruction count.  Everything else follows conv_halo.cu: ONE TMA halo load per 64-channel chunk, views
// addressed through the UMMA descriptor, weights resident in shared memory (RC) or streamed, persistent CTAs,
// double-buffered TMEM accumulators, PDL prologue overlap, fused bias / residual / ReLU / 1x1 head + sigmoid epilogue.
#include <cuda.h>

#include <atomic>

#include "conv_halo.h"
#include "ltb_internal.h"
#include "ptx_sm100.cuh"

namespace ltb {

constexpr int kYsP = 10;  // halo row pitch in pixels (8 + 2)

template <int BN, int NSUB, int RC>
struct YsCfg {
  static constexpr int HR = 16 * NSUB;                          // halo rows = MMA rows / 8 (no +-1 rows: the y shift happens in the epilogue)
  static constexpr int A_BYTES_RAW = HR * kYsP * 128;
  static constexpr int A_BYTES = (A_BYTES_RAW + 1023) & ~1023;
  static constexpr int B_BYTES = 3 * BN * 128;                  // taps (dy = 0,1,2) of one dx, BN rows each, 64 k
  static constexpr int XBUF_BYTES = 8 * 2 * 32 * 8 * 4;         // 8 epilogue warps x {bottom row group of D_0, top row group of D_2} x 32 cols x 8 lanes
  static constexpr int BUDGET = 226 * 1024 - XBUF_BYTES;
  static constexpr int A_FIT = (BUDGET - RC * 3 * B_BYTES) / A_BYTES;
  static constexpr int A_STAGES = RC ? (A_FIT > 4 ? 4 : A_FIT) : 3;
  static constexpr int B_FIT = (BUDGET - A_STAGES * A_BYTES) / B_BYTES;
  static constexpr int B_STAGES = RC ? RC * 3 : (B_FIT > 6 ? 6 : B_FIT);
  static constexpr int ACC_COLS = NSUB * 3 * BN;
  static constexpr int TCOLS = (2 * ACC_COLS <= 256) ? 256 : 512;
  static constexpr int NCHO = NSUB * BN / 32;                   // 32-channel output items per tile: one per epilogue warp group
  static constexpr int SMEM_BYTES = A_STAGES * A_BYTES + B_STAGES * B_BYTES + XBUF_BYTES + 1024;
  static_assert(2 * ACC_COLS <= 512, "TMEM overflow");
  static_assert(3 * BN <= 256 && (3 * BN) % 16 == 0, "stacked N must be a legal UMMA N");
  static_assert(NCHO >= 1 && NCHO <= 2, "one output item per warp group: (BN,NSUB) in {(64,1),(32,2),(32,1)}");
  static_assert(A_STAGES >= 2 && B_STAGES >= 3, "shared memory budget");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory overflow");
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int BN, int NSUB, int RC>
__global__ void __launch_bounds__(320, 1) conv_ystack_umma_kernel(const __grid_constant__ HaloParams p) {
  using C = YsCfg<BN, NSUB, RC>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[C::A_STAGES], a_empty[C::A_STAGES];
  __shared__ __align__(8) uint64_t b_full[C::B_STAGES], b_empty[C::B_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float head_sw[100];   // fused head: 3 x 32 weights + 3 biases

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr bool kHeadOk = (BN == 32);
  if (kHeadOk && p.head_out && tid >= 64 && tid < 64 + 99) head_sw[tid - 64] = (tid - 64 < 96) ? p.head_w[tid - 64] : p.head_b[tid - 64 - 96];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = smem0;
  const uint32_t b_smem = smem0 + C::A_STAGES * C::A_BYTES;
  float* const xbuf = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)) + C::A_STAGES * C::A_BYTES + C::B_STAGES * C::B_BYTES);
  const int chunks = (p.Cin + 63) / 64;

  if (tid == 0) {
    for (int s = 0; s < C::A_STAGES; ++s) {
      mbar_init(smem_u32(&a_full[s]), 1);
      mbar_init(smem_u32(&a_empty[s]), 1);
    }
    for (int s = 0; s < C::B_STAGES; ++s) {
      mbar_init(smem_u32(&b_full[s]), 1);
      mbar_init(smem_u32(&b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&acc_full[s]), 1);
      mbar_init(smem_u32(&acc_empty[s]), 8);
    }
    mbar_fence_init();
    tma_prefetch_desc(&p.tm_in);
    tma_prefetch_desc(&p.tm_w);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), C::TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int tiles_img = p.tiles_x * p.tiles_y;

  pdl_launch_dependents();
  if (RC && warp == 0 && lane == 0) {
    // resident weights (constants): fetched before this kernel waits for its predecessor
    mbar_arrive_expect_tx(smem_u32(&b_full[0]), RC * 3 * C::B_BYTES);
    for (int c = 0; c < RC; ++c)
      for (int j = 0; j < 3; ++j) tma_load_4d(b_smem + (c * 3 + j) * C::B_BYTES, &p.tm_w, smem_u32(&b_full[0]), c * 64, 0, j, 0);
  }
  pdl_wait();

  if (warp == 0) {
    // =============================================================== TMA producer
    if (lane == 0) {
      uint32_t ai = 0, bi = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int img = t / tiles_img;
        const int mt = t - img * tiles_img;
        const int ty = mt / p.tiles_x, tx = mt - ty * p.tiles_x;
        const int y0 = ty * p.tile_rows - 1, x0 = tx * 8 - 1;   // MMA row 0 = output row y0 (only its D_0 is used), halo column 0 = x0
        for (int c = 0; c < chunks; ++c) {
          const uint32_t as = ai % C::A_STAGES;
          mbar_wait(smem_u32(&a_empty[as]), ((ai / C::A_STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(smem_u32(&a_full[as]), C::A_BYTES_RAW);
          tma_load_4d(a_smem + as * C::A_BYTES, &p.tm_in, smem_u32(&a_full[as]), c * 64, x0, y0, img);
          ++ai;
          if (RC) continue;
          for (int j = 0; j < 3; ++j) {
            const uint32_t bs = bi % C::B_STAGES;
            mbar_wait(smem_u32(&b_empty[bs]), ((bi / C::B_STAGES) & 1u) ^ 1u);
            mbar_arrive_expect_tx(smem_u32(&b_full[bs]), C::B_BYTES);
            tma_load_4d(b_smem + bs * C::B_BYTES, &p.tm_w, smem_u32(&b_full[bs]), c * 64, 0, j, 0);
            ++bi;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer (warp-uniform, elected-lane predication)
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc = umma_idesc_f16(128, 3 * BN);
    constexpr uint32_t kADescHi = ((kYsP * 128) >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t kBDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    uint32_t ai = 0, bi = 0, it = 0;
    if (RC) mbar_wait(smem_u32(&b_full[0]), 0);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      mbar_wait(smem_u32(&acc_empty[buf]), ((it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t dbase = tmem + buf * C::ACC_COLS;
#pragma unroll 1
      for (int c = 0; c < chunks; ++c) {
        const uint32_t as = ai % C::A_STAGES;
        mbar_wait(smem_u32(&a_full[as]), (ai / C::A_STAGES) & 1u);
        const uint32_t a_lo0 = (((a_smem + as * C::A_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
        const int ksteps = (c == chunks - 1) ? p.last_ksteps : 4;
        const uint32_t later = (c != 0) ? 1u : 0u;
#pragma unroll
        for (int j = 0; j < 3; ++j) {   // j = dx: the A view is the halo shifted by j pixels, the B stage holds W[0..2][j]
          const uint32_t bs = RC ? (uint32_t)(c * 3 + j) : bi % C::B_STAGES;
          if (!RC) mbar_wait(smem_u32(&b_full[bs]), (bi / C::B_STAGES) & 1u);
          tc_fence_after();
          const uint32_t b_lo = (((b_smem + bs * C::B_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t acc0 = later | (j ? 1u : 0u);
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub) {
            const uint32_t d = dbase + sub * (3 * BN);
            const uint32_t a_lo = a_lo0 + (uint32_t)j * 8u + sub * (16 * kYsP * 8);
            if (ksteps == 4) {
              umma_f16_lohi_x4_if(leader, d, a_lo, kADescHi, b_lo, kBDescHi, idesc, acc0);
            } else {
#pragma unroll 1
              for (int k = 0; k < ksteps; ++k)
                umma_f16_lohi_if(leader, d, a_lo + k * 2, kADescHi, b_lo + k * 2, kBDescHi, idesc, k ? 1u : acc0);
            }
          }
          if (!RC) umma_commit_if(leader, smem_u32(&b_empty[bs]));
          ++bi;
        }
        umma_commit_if(leader, smem_u32(&a_empty[as]));
        ++ai;
      }
      umma_commit_if(leader, smem_u32(&acc_full[buf]));
    }
    __syncwarp();
  } else {
    // =============================================================== epilogue: warps 2..9, TMEM lane quarter q = warp % 4,
    // group grp = 0/1 owns output item ci = grp: (sub, c0) = (ci*32 / BN, ci*32 % BN).  wid = grp*4 + q orders the warps so
    // that wid +- 1 is the neighbouring row quarter of the SAME 32-channel item (also across the sub-tile boundary).
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int wid = grp * 4 + q;
    const bool active = grp < C::NCHO;
    const int sub = active ? (grp * 32) / BN : 0;
    const int c0 = active ? (grp * 32) % BN : 0;
    const int Q = sub * 4 + q;                        // row quarter inside the tile: rows 4Q .. 4Q+3
    const int R = Q * 4 + (lane >> 3), rx = lane & 7;  // tile-local MMA row, column
    const bool has_up = Q > 0, has_dn = Q < 4 * NSUB - 1;
    float* const xb_bot = xbuf + (size_t)wid * 512;        // D_0 of this quarter's LAST row group   [32 cols][8 lanes]
    float* const xb_top = xbuf + (size_t)wid * 512 + 256;  // D_2 of this quarter's FIRST row group
    const bool has_res = p.res != nullptr;
    const bool head = kHeadOk && p.head_out != nullptr;

    auto tile_pix = [&](int t, size_t& opix, bool& ok) {
      const int img = t / tiles_img;
      const int mt = t - img * tiles_img;
      const int ty = mt / p.tiles_x, tx = mt - ty * p.tiles_x;
      const int gy = ty * p.tile_rows + R - 1, gx = tx * 8 + rx;
      ok = active && R >= 1 && R <= 16 * NSUB - 2 && gy < p.GH && gx < p.GW;
      opix = ok ? ((size_t)img * p.OH + gy) * p.OW + gx : 0;
    };
    // residual: 64 B per thread, requested one whole TILE ahead (the narrow layers were latency bound on this load)
    uint4 rnext[4];
    auto load_res = [&](int t) {
      size_t opix;
      bool ok;
      tile_pix(t, opix, ok);
      if (ok) {
        const __half* rptr = p.res + opix * p.RCtot + p.rc_off + c0;
        if (p.wide_io) {
          ldg256(rptr, rnext[0], rnext[1]);
          ldg256(rptr + 16, rnext[2], rnext[3]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) rnext[u] = __ldcg(reinterpret_cast<const uint4*>(rptr + 8 * u));
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) rnext[u] = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    if (has_res && (int)blockIdx.x < p.total_tiles) load_res(blockIdx.x);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      size_t opix;
      bool row_ok;
      tile_pix(t, opix, row_ok);
      uint4 rcur[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) rcur[u] = rnext[u];
      if (has_res && t + (int)gridDim.x < p.total_tiles) load_res(t + gridDim.x);
      mbar_wait(smem_u32(&acc_full[buf]), (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t tb = tmem + buf * C::ACC_COLS + ((uint32_t)(q * 32) << 16) + sub * (3 * BN) + c0;
      uint32_t v0[32], v1[32], v2[32];
      if (active) {
        tmem_ld32(tb, v0);
        tmem_ld32(tb + 2 * BN, v2);
        tmem_ld32(tb + BN, v1);
      }
      tmem_ld_wait();
      // the accumulator buffer is free as soon as its columns are in registers: the MMAs of tile it+2 may start
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
      // ---- exchange the boundary row groups between neighbouring quarters
      named_bar_sync(1, 256);                         // every warp has finished reading the previous tile's exchange data
      if (active) {
        if (lane >= 24) {
#pragma unroll
          for (int c = 0; c < 32; ++c) xb_bot[c * 8 + (lane - 24)] = __uint_as_float(v0[c]);
        }
        if (lane < 8) {
#pragma unroll
          for (int c = 0; c < 32; ++c) xb_top[c * 8 + lane] = __uint_as_float(v2[c]);
        }
      }
      named_bar_sync(1, 256);
      if (!active) continue;
      const float* up_src = xb_bot - 512;             // quarter Q-1 (warp wid-1): D_0 of its last row group
      const float* dn_src = xb_top + 512;             // quarter Q+1 (warp wid+1): D_2 of its first row group
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float up = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[c]), 8);
        float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[c]), 8);
        if (lane < 8) up = has_up ? up_src[c * 8 + lane] : 0.f;
        if (lane >= 24) dn = has_dn ? dn_src[c * 8 + (lane - 24)] : 0.f;
        v1[c] = __float_as_uint(__uint_as_float(v1[c]) + up + dn);
      }
      // ---- bias / residual / ReLU / (head) / store: same arithmetic and order as conv_halo.cu's epilogue
      float4 bb[8];   // (hoisting these 32 registers out of the tile loop spills: 168-register cap at 320 threads)
#pragma unroll
      for (int u = 0; u < 8; ++u) bb[u] = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + u);
      __half* optr = p.out + opix * p.OCtot + p.oc_off;
      float ha0 = 0.f, ha1 = 0.f, ha2 = 0.f;
      if (head) {
        ha0 = head_sw[96];
        ha1 = head_sw[97];
        ha2 = head_sw[98];
      }
#pragma unroll
      for (int g16 = 0; g16 < 32; g16 += 16) {
        uint4 ovv[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int g = g16 + hh * 8;
          const float4 b0 = bb[g / 4], b1 = bb[g / 4 + 1];
          __half2* oh = reinterpret_cast<__half2*>(&ovv[hh]);
          uint32_t* ow = reinterpret_cast<uint32_t*>(&ovv[hh]);
          const float f0 = __uint_as_float(v1[g + 0]) + b0.x, f1 = __uint_as_float(v1[g + 1]) + b0.y;
          const float f2 = __uint_as_float(v1[g + 2]) + b0.z, f3 = __uint_as_float(v1[g + 3]) + b0.w;
          const float f4 = __uint_as_float(v1[g + 4]) + b1.x, f5 = __uint_as_float(v1[g + 5]) + b1.y;
          const float f6 = __uint_as_float(v1[g + 6]) + b1.z, f7 = __uint_as_float(v1[g + 7]) + b1.w;
          if (!has_res && p.relu) {
            ow[0] = f32x2_to_f16x2_sat_relu(f0, f1);
            ow[1] = f32x2_to_f16x2_sat_relu(f2, f3);
            ow[2] = f32x2_to_f16x2_sat_relu(f4, f5);
            ow[3] = f32x2_to_f16x2_sat_relu(f6, f7);
          } else {
            ow[0] = f32x2_to_f16x2_sat(f0, f1);
            ow[1] = f32x2_to_f16x2_sat(f2, f3);
            ow[2] = f32x2_to_f16x2_sat(f4, f5);
            ow[3] = f32x2_to_f16x2_sat(f6, f7);
            if (has_res) {
              const __half2* rh = reinterpret_cast<const __half2*>(&rcur[(g16 >> 3) + hh]);
              const __half2 hmax = __floats2half2_rn(65504.f, 65504.f);
              const __half2 lo = p.relu ? __floats2half2_rn(0.f, 0.f) : __floats2half2_rn(-65504.f, -65504.f);
#pragma unroll
              for (int u = 0; u < 4; ++u) oh[u] = __hmin2(__hmax2(__hadd2(oh[u], rh[u]), lo), hmax);
            }
          }
          if (kHeadOk && head) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 f = __half22float2(oh[u]);
              const int c = g + 2 * u;
              ha0 = fmaf(f.x, head_sw[c], ha0);
              ha0 = fmaf(f.y, head_sw[c + 1], ha0);
              ha1 = fmaf(f.x, head_sw[32 + c], ha1);
              ha1 = fmaf(f.y, head_sw[32 + c + 1], ha1);
              ha2 = fmaf(f.x, head_sw[64 + c], ha2);
              ha2 = fmaf(f.y, head_sw[64 + c + 1], ha2);
            }
          }
        }
        if (row_ok && !head) {
          if (p.wide_io) {
            stg256(optr + c0 + g16, ovv[0], ovv[1]);
          } else {
            *reinterpret_cast<uint4*>(optr + c0 + g16) = ovv[0];
            *reinterpret_cast<uint4*>(optr + c0 + g16 + 8) = ovv[1];
          }
        }
      }
      if (kHeadOk && head && row_ok) {
        float* o = p.head_out + opix * 3;
        o[0] = (1.f / (1.f + expf(-ha0))) * 255.f;
        o[1] = (1.f / (1.f + expf(-ha1))) * 255.f;
        o[2] = (1.f / (1.f + expf(-ha2))) * 255.f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, C::TCOLS);
  }
}

template <int BN, int NSUB, int RC>
static cudaError_t launch_ys(const HaloPlan& pl, int sms, cudaStream_t st) {
  using C = YsCfg<BN, NSUB, RC>;
  static SmemConfigOnce once;
  if (cudaError_t e = once.ensure(conv_ystack_umma_kernel<BN, NSUB, RC>, C::SMEM_BYTES); e != cudaSuccess) return e;
  const int grid = pl.hp.total_tiles < sms ? pl.hp.total_tiles : sms;
  return launch_kernel_pdl(conv_ystack_umma_kernel<BN, NSUB, RC>, dim3(grid), dim3(320), C::SMEM_BYTES, st, pl.hp);
}

cudaError_t launch_conv_ystack(const HaloPlan& pl, int sms, cudaStream_t st) {
  const int chunks = (pl.hp.Cin + 63) / 64;
  const bool resident = pl.hp.total_tiles >= 2 * sms && chunks <= 2;   // every tile uses the same (whole) weight set
  if (pl.BN == 64 && pl.NSUB == 1) {
    if (resident && chunks == 1) return launch_ys<64, 1, 1>(pl, sms, st);
    return launch_ys<64, 1, 0>(pl, sms, st);
  }
  if (pl.BN == 32 && pl.NSUB == 2) {
    if (resident) return chunks == 1 ? launch_ys<32, 2, 1>(pl, sms, st) : launch_ys<32, 2, 2>(pl, sms, st);
    return launch_ys<32, 2, 0>(pl, sms, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace ltb
