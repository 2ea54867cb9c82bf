"""The reference's frame watermark for frames that never leave the device (output side).

``BaseAvatar.process_frames`` draws ``cv2.putText(frame, "LiveTalking", (10, 20), cv2.FONT_HERSHEY_SIMPLEX, 0.3, (128,128,128), 1)``
into every frame right before it is pushed to the output (avatars/base_avatar.py:449).  With thickness 1 and the default LINE_8
OpenCV sets a frame-independent set of pixels to the colour, without blending — so OpenCV itself rasterises the text once, here,
into a scratch image, and the resulting pixel list is stamped on the device (``Ctx.stamp_pixels`` -> ``ltb_op_stamp_pixels``) before
the BGR -> I420 conversion of the encoder hand-off.  Bit-exact with the reference by construction (tests/test_watermark.py)."""
from __future__ import annotations

import numpy as np

TEXT, ORG, SCALE, COLOR, THICKNESS = "LiveTalking", (10, 20), 0.3, (128, 128, 128), 1


def text_pixels(H: int, W: int, text: str = TEXT, org=ORG, scale: float = SCALE, thickness: int = THICKNESS) -> np.ndarray:
    """(n, 2) int32 (y, x) — the pixels cv2.putText writes for this text on an H x W frame (clipped to the frame like OpenCV does)."""
    import cv2
    probe = np.zeros((H, W, 3), np.uint8)
    cv2.putText(probe, text, tuple(org), cv2.FONT_HERSHEY_SIMPLEX, scale, (255, 255, 255), thickness)
    ys, xs = np.nonzero(probe[..., 0])
    return np.ascontiguousarray(np.stack([ys, xs], 1).astype(np.int32))


def stamp_host(frame: np.ndarray, pixels: np.ndarray, color=COLOR) -> np.ndarray:
    """numpy form of the device stamp (oracle for the GPU test; in place)."""
    frame[pixels[:, 0], pixels[:, 1]] = np.asarray(color, np.uint8)
    return frame
