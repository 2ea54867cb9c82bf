// Hardware probe (test hook, not on the product path): how does tcgen05.mma read a K-major SWIZZLE_128B operand whose
// start address is NOT aligned to the 1024-byte swizzle repeat, and whose 8-row groups are SBO bytes apart with SBO not a
// multiple of 1024?  The halo-resident 3x3 conv (conv_halo.cu) relies on the answer: it keeps ONE (TH+2)x(TW+2) input halo
// in shared memory and addresses the nine shifted im2col views of it purely through descriptor start/SBO fields.
#include "../ltb_internal.h"
#include "../ptx_sm100.cuh"
#include "../../../include/ltb200_diag.h"

namespace ltb {

__global__ void __launch_bounds__(128) umma_probe_kernel(const __half* __restrict__ halo, int halo_rows, const __half* __restrict__ bmat,
                                                         int start_row, int sbo_rows, int base_offset, float* __restrict__ out, int noswz_lbo_bytes, int noswz_sbo_bytes,
                                                         int noswz_start_bytes) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;                                   // halo_rows x 128 B
  const uint32_t b_base = base + ((halo_rows * 128 + 1023) & ~1023);  // 64 x 128 B
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  for (int q = tid; q < halo_rows * 8; q += 128) {
    const int r = q >> 3, j = q & 7;
    const int dst = noswz_lbo_bytes ? (r * 128 + (j << 4)) : (r * 128 + ((j ^ (r & 7)) << 4));   // linear copy in no-swizzle mode
    *reinterpret_cast<uint4*>(gen + dst) = *reinterpret_cast<const uint4*>(halo + (size_t)r * 64 + j * 8);
  }
  uint8_t* genb = gen + (b_base - base);
  for (int q = tid; q < 64 * 8; q += 128) {
    const int r = q >> 3, j = q & 7;
    *reinterpret_cast<uint4*>(genb + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(bmat + (size_t)r * 64 + j * 8);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 64);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128, 64);
    for (int k = 0; k < 4; ++k) {
      uint64_t ad;
      if (noswz_lbo_bytes) {
        // SWIZZLE_NONE K-major: core matrix = 8 rows x 16 B (rows 16 B apart), K halves LBO apart, 8-row groups SBO apart
        const uint32_t sa = a_base + noswz_start_bytes + k * 2 * noswz_lbo_bytes;
        ad = (uint64_t)((sa & 0x3FFFF) >> 4) | ((uint64_t)((noswz_lbo_bytes >> 4) & 0x3FFF) << 16) |
             ((uint64_t)((noswz_sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
      } else {
        ad = umma_smem_desc(a_base + start_row * 128 + k * 32, sbo_rows * 128, 2);
        ad |= (uint64_t)(base_offset & 7) << 49;
      }
      const uint64_t bd = umma_smem_desc(b_base + k * 32, 1024, 2);
      umma_f16(tmem, ad, bd, idesc, k != 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(size_t)tid * 64 + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

}  // namespace ltb

extern "C" int ltb_umma_probe_noswz(const void* buf_f16, int buf_rows, const void* b_f16, int start_bytes, int lbo_bytes, int sbo_bytes,
                                    float* out_128x64);

static int probe_run(const void* halo_f16, int halo_rows, const void* b_f16, int start_row, int sbo_rows, int base_offset, float* out_128x64,
                     int lbo_b, int sbo_b, int start_b);

extern "C" int ltb_umma_probe(const void* halo_f16, int halo_rows, const void* b_f16, int start_row, int sbo_rows, int base_offset,
                              float* out_128x64) {
  if (halo_rows < start_row + 15 * sbo_rows + 8) return ltb::fail(__FILE__, __LINE__, "probe: halo too small");
  return probe_run(halo_f16, halo_rows, b_f16, start_row, sbo_rows, base_offset, out_128x64, 0, 0, 0);
}

extern "C" int ltb_umma_probe_noswz(const void* buf_f16, int buf_rows, const void* b_f16, int start_bytes, int lbo_bytes, int sbo_bytes,
                                    float* out_128x64) {
  if (lbo_bytes < 16 || (lbo_bytes & 15) || (sbo_bytes & 15) || (start_bytes & 15)) return ltb::fail(__FILE__, __LINE__, "probe: bad no-swizzle strides");
  if ((long)start_bytes + 15L * sbo_bytes + 7 * 16 + 7L * lbo_bytes + 16 > (long)buf_rows * 128) return ltb::fail(__FILE__, __LINE__, "probe: buffer too small");
  return probe_run(buf_f16, buf_rows, b_f16, 0, 8, 0, out_128x64, lbo_bytes, sbo_bytes, start_bytes);
}

static int probe_run(const void* halo_f16, int halo_rows, const void* b_f16, int start_row, int sbo_rows, int base_offset, float* out_128x64,
                     int lbo_b, int sbo_b, int start_b) {
  using namespace ltb;
  if (!halo_f16 || !b_f16 || !out_128x64) return LTB_FAIL("null argument");
  if (halo_rows > 1500) return LTB_FAIL("probe: halo too large");
  __half *dh = nullptr, *db = nullptr;
  float* dout = nullptr;
  const int smem = ((halo_rows * 128 + 1023) & ~1023) + 8192 + 1024;
  LTB_CUDA(cudaMalloc(&dh, (size_t)halo_rows * 128));
  LTB_CUDA(cudaMalloc(&db, 8192));
  LTB_CUDA(cudaMalloc(&dout, 128 * 64 * 4));
  LTB_CUDA(cudaMemcpy(dh, halo_f16, (size_t)halo_rows * 128, cudaMemcpyHostToDevice));
  LTB_CUDA(cudaMemcpy(db, b_f16, 8192, cudaMemcpyHostToDevice));
  LTB_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe_kernel<<<1, 128, smem>>>(dh, halo_rows, db, start_row, sbo_rows, base_offset, dout, lbo_b, sbo_b, start_b);
  LTB_CUDA(cudaGetLastError());
  LTB_CUDA(cudaDeviceSynchronize());
  LTB_CUDA(cudaMemcpy(out_128x64, dout, 128 * 64 * 4, cudaMemcpyDeviceToHost));
  cudaFree(dh);
  cudaFree(db);
  cudaFree(dout);
  return 0;
}
