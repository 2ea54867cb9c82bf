"""CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py): BGR -> planar YUV 4:2:0 (I420) for the encoder hand-off.

SURVEY.md §8(f) rank 3.  The reference hands every composited BGR frame to PyAV (``VideoFrame.from_ndarray(frame,
format="bgr24")``, avatars/base_avatar.py:449-453) and aiortc's H.264 encoder converts it to yuv420p on the CPU inside
libswscale (server/webrtc.py) — a third-party dependency that is ABSENT from this image (``import av`` fails), so its
exact arithmetic cannot be pinned here: **parity unpinned against libswscale**.  What IS pinned, bit for bit, is OpenCV's
``cv2.cvtColor(img, cv2.COLOR_BGR2YUV_I420)`` (cv2 4.13 is installed), restated below from imgproc/color_yuv: BT.601
limited range, 20-bit fixed point, chroma taken from the top-left pixel of every 2x2 block (no averaging).  Layout of the
result: (H*3/2, W) uint8 = Y plane (H x W), then U (H/2 x W/2), then V (H/2 x W/2), each contiguous.
"""
from __future__ import annotations

import numpy as np

SHIFT = 20
CRY, CGY, CBY = 269484, 528482, 102760
CRU, CGU, CBU = -155188, -305135, 460324
CGV, CBV = -385875, -74448
HALF = 1 << (SHIFT - 1)


def bgr_to_i420(img: np.ndarray) -> np.ndarray:
    """img: (H, W, 3) uint8 BGR with even H and W -> (H*3//2, W) uint8, as cv2.COLOR_BGR2YUV_I420."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
    H, W = img.shape[:2]
    assert H % 2 == 0 and W % 2 == 0, "I420 needs even dimensions"
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    y = (CRY * r + CGY * g + CBY * b + HALF + (16 << SHIFT)) >> SHIFT
    r0, g0, b0 = r[0::2, 0::2], g[0::2, 0::2], b[0::2, 0::2]
    u = (CRU * r0 + CGU * g0 + CBU * b0 + HALF + (128 << SHIFT)) >> SHIFT
    v = (CBU * r0 + CGV * g0 + CBV * b0 + HALF + (128 << SHIFT)) >> SHIFT
    out = np.empty((H * 3 // 2, W), np.uint8)
    out[:H] = np.clip(y, 0, 255)
    out[H:].reshape(-1)[: (H // 2) * (W // 2)] = np.clip(u, 0, 255).reshape(-1)
    out[H:].reshape(-1)[(H // 2) * (W // 2):] = np.clip(v, 0, 255).reshape(-1)
    return out
