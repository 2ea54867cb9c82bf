"""Hardware probe (run manually on the GPU box): unaligned-start / odd-SBO behaviour of tcgen05 smem descriptors.
Needs the diagnostic library: `python -m livetalking_b200.build --diag` (the probes are not part of libltb200.so)."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from livetalking_b200 import _capi
_capi.LIB_PATH = os.path.join(os.path.dirname(_capi.LIB_PATH), "libltb200_diag.so")
from livetalking_b200 import engine

engine.set_device(0)
lib = _capi.lib()
rng = np.random.default_rng(0)
R = 400
halo = (rng.standard_normal((R, 64)) * 0.5).astype(np.float16)
b = (rng.standard_normal((64, 64)) * 0.5).astype(np.float16)
out = np.zeros((128, 64), np.float32)
for start, sbo, bo in [(0, 8, 0), (8, 8, 0), (3, 8, 3), (3, 8, 0), (11, 10, 3), (11, 10, 0), (0, 10, 0), (16, 10, 0), (1, 16, 1), (1, 16, 0),
                       (5, 16, 5), (5, 16, 0), (21, 10, 5), (21, 10, 0), (2, 18, 2), (2, 18, 0)]:
    _capi.check(lib.ltb_umma_probe(halo.ctypes.data, R, b.ctypes.data, start, sbo, bo, out.ctypes.data))
    rows = np.array([start + (m // 8) * sbo + m % 8 for m in range(128)])
    ref = halo[rows].astype(np.float32) @ b.astype(np.float32).T
    err = np.abs(out - ref).max()
    # which groups are right?
    good = [int(np.abs(out[g * 8:(g + 1) * 8] - ref[g * 8:(g + 1) * 8]).max() < 1e-2) for g in range(16)]
    print(f"start={start:3d} sbo_rows={sbo:3d} base_offset={bo}: max err {err:.4f} groups_ok={''.join(map(str, good))}")

print("---- SWIZZLE_NONE probes (overlapping rows)")
buf = (rng.standard_normal((R, 64)) * 0.5).astype(np.float16)
flat = buf.reshape(-1)
for start_b, lbo, sbo in [(0, 128, 1024), (0, 16, 256), (48, 16, 256), (256 * 3 + 32, 16, 256), (0, 16, 128), (16, 32, 512)]:
    rc = lib.ltb_umma_probe_noswz(buf.ctypes.data, R, b.ctypes.data, start_b, lbo, sbo, out.ctypes.data)
    if rc != 0:
        print("probe refused:", lib.ltb_last_error().decode()); continue
    A = np.zeros((128, 64), np.float32)
    for m in range(128):
        for j in range(8):
            off = (start_b + (m // 8) * sbo + (m % 8) * 16 + j * lbo) // 2
            A[m, j * 8:(j + 1) * 8] = flat[off:off + 8].astype(np.float32)
    ref = A @ b.astype(np.float32).T
    print(f"noswz start={start_b} lbo={lbo} sbo={sbo}: max err {np.abs(out - ref).max():.4f}")
