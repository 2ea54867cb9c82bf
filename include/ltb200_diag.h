/*
 * libltb200_diag.so only — hardware probes used while developing the tcgen05 kernels (tools/diag_halo.py, tests/probe_umma.py).
 * NOT part of the product library: livetalking_b200/build.py compiles csrc/diag/*.cu only into the diagnostic build
 * (python -m livetalking_b200.build --diag).
 */
#ifndef LTB200_DIAG_H_
#define LTB200_DIAG_H_
#ifdef __cplusplus
extern "C" {
#endif

/* hardware probe (test hook): D[128x64] = A * B^T with A = 16 groups of 8 consecutive 128-byte rows of a swizzled
 * shared-memory buffer, first group at row `start_row`, groups `sbo_rows` rows apart, descriptor base_offset as given. */
int ltb_umma_probe(const void* halo_f16, int halo_rows, const void* b_f16, int start_row, int sbo_rows, int base_offset,
                   float* out_128x64);

/* same probe for SWIZZLE_NONE K-major operands: buffer copied linearly; row m, k-chunk j (8 halves) of A is read at
 * start_bytes + (m/8)*sbo_bytes + (m%8)*16 + j*lbo_bytes — lbo_bytes = 16 makes consecutive rows overlap (im2col of an
 * 8-channel image without materialising it). */
int ltb_umma_probe_noswz(const void* buf_f16, int buf_rows, const void* b_f16, int start_bytes, int lbo_bytes, int sbo_bytes,
                         float* out_128x64);

#ifdef __cplusplus
}
#endif
#endif
