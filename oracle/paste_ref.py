"""CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py): paste-back composite.

numpy restatement of

    avatars/wav2lip_avatar.py:141-147   LipReal.paste_back_frame
        bbox = (y1, y2, x1, x2); frame.copy(); pred.astype(uint8) (TRUNCATION);
        cv2.resize(..., (x2-x1, y2-y1))  (INTER_LINEAR default);  overwrite rectangle.

``cv2.resize`` is OpenCV (third-party; cv2 4.13 in this image).  Its 8-bit INTER_LINEAR
path is fixed-point: 11-bit coefficients (INTER_RESIZE_COEF_BITS = 11), int32 horizontal
pass, and the vertical pass  ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2 ; an exact
2x decimation in both axes is silently switched to INTER_AREA (2x2 box average with
round-half-up).  Restated from OpenCV's published imgproc/resize.cpp algorithm and pinned
BIT-EXACT against the installed cv2 by tests/test_oracle_paste.py over randomized sizes.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _linear_coeffs(dst: int, src: int):
    """Per-destination-index source offset and 2 fixed-point taps (cv::resize, linear, 8U)."""
    scale = 1.0 / (float(dst) / float(src))          # double, as resize.cpp: scale_x = 1./inv_scale_x
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)  # fx = (float)((dx+0.5)*scale_x - 0.5)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def _coef_short(v32: np.ndarray) -> np.ndarray:
    # saturate_cast<short>(float * 2048): cvRound = round-half-to-even
    return np.clip(np.rint(v32.astype(np.float32) * np.float32(COEF_SCALE)), -32768, 32767).astype(np.int64)


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src_u8_HxWxC, (dw, dh)) with the default INTER_LINEAR, bit-exact."""
    src = np.ascontiguousarray(src)
    assert src.dtype == np.uint8 and src.ndim == 3
    sh, sw, cn = src.shape
    if (dw, dh) == (sw, sh):
        return src.copy()
    if sw == 2 * dw and sh == 2 * dh:
        # INTER_LINEAR + exact 2x decimation -> INTER_AREA fast path: (a+b+c+d+2)>>2
        s = src.astype(np.int64)
        acc = s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2]
        return ((acc + 2) >> 2).astype(np.uint8)

    sx, fx = _linear_coeffs(dw, sw)
    # horizontal clamping (resize.cpp: sx<0 -> sx=0,fx=0 ; sx>=w-1 -> sx=w-1,fx=0)
    lo = sx < 0
    sx = np.where(lo, 0, sx)
    fx = np.where(lo, np.float32(0), fx)
    hi = sx >= sw - 1
    sx = np.where(hi, sw - 1, sx)
    fx = np.where(hi, np.float32(0), fx).astype(np.float32)
    a0 = _coef_short(np.float32(1.0) - fx)
    a1 = _coef_short(fx)
    sx1 = np.minimum(sx + 1, sw - 1)          # tap 1 has weight 0 whenever it would be out of range

    sy, fy = _linear_coeffs(dh, sh)
    b0 = _coef_short(np.float32(1.0) - fy)
    b1 = _coef_short(fy)
    sy0 = np.clip(sy, 0, sh - 1)
    sy1 = np.clip(sy + 1, 0, sh - 1)

    s = src.astype(np.int64)
    # horizontal pass on every source row: (sh, dw, cn) int32 values
    hrow = s[:, sx, :] * a0[None, :, None] + s[:, sx1, :] * a1[None, :, None]
    S0 = hrow[sy0]
    S1 = hrow[sy1]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def w2l_paste_back(pred_f32: np.ndarray, frame_u8: np.ndarray, bbox) -> np.ndarray:
    """wav2lip_avatar.py:141-147.  pred (256,256,3) float in [0,255]; bbox = (y1,y2,x1,x2)."""
    y1, y2, x1, x2 = [int(v) for v in bbox]
    out = frame_u8.copy()
    face = np.asarray(pred_f32).astype(np.uint8)      # truncation toward zero, as the reference
    out[y1:y2, x1:x2] = resize_linear_u8(face, x2 - x1, y2 - y1)
    return out


def mirror_index(size: int, index: int) -> int:
    """utils/image.py:26-32 — ping-pong index over the avatar clip."""
    turn = index // size
    res = index % size
    return res if turn % 2 == 0 else size - res - 1


def w2l_build_batch(face_list, index: int, batch: int):
    """wav2lip_avatar.py:116-130: gather faces by mirror index, zero the lower half of the
    masked copy, concat (masked, full) on channels, /255 (float64) -> (B,6,256,256) float32."""
    length = len(face_list)
    faces = np.asarray([face_list[mirror_index(length, index + i)] for i in range(batch)])
    masked = faces.copy()
    masked[:, faces.shape[1] // 2:] = 0
    img = np.concatenate((masked, faces), axis=3) / 255.0
    return np.transpose(img, (0, 3, 1, 2)).astype(np.float32)


# --------------------------------------------------------------------------------------------------------------------
# MuseTalk blend paste-back
# --------------------------------------------------------------------------------------------------------------------
def bgr2gray_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) for uint8: 15-bit fixed point, pinned against the installed cv2."""
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


def blend_linear_u8(src1: np.ndarray, src2: np.ndarray, w1: np.ndarray, w2: np.ndarray) -> np.ndarray:
    """cv2.blendLinear for uint8 images / float32 weights: sat_u8(rint((s1*w1 + s2*w2) / (w1 + w2 + 1e-5))) in float32."""
    w1 = w1.astype(np.float32)[..., None]
    w2 = w2.astype(np.float32)[..., None]
    den = (w1 + w2 + np.float32(1e-5)).astype(np.float32)
    num = (src1.astype(np.float32) * w1).astype(np.float32) + (src2.astype(np.float32) * w2).astype(np.float32)
    return np.clip(np.rint((num / den).astype(np.float32)), 0, 255).astype(np.uint8)


def mt_paste_back(pred_u8: np.ndarray, frame_u8: np.ndarray, bbox, mask_u8: np.ndarray, crop_box) -> np.ndarray:
    """MuseReal.paste_back_frame (avatars/musetalk_avatar.py:154-164) + get_image_blending (avatars/musetalk/myutil.py:4-25).
    bbox = (x1,y1,x2,y2); crop_box = (x_s,y_s,x_e,y_e); mask: (y_e-y_s, x_e-x_s, 3) uint8."""
    x1, y1, x2, y2 = [int(v) for v in bbox]
    xs, ys, xe, ye = [int(v) for v in crop_box]
    body = frame_u8.copy()
    res = resize_linear_u8(np.asarray(pred_u8).astype(np.uint8), x2 - x1, y2 - y1)
    face_large = body[ys:ye, xs:xe].copy()
    face_large[y1 - ys:y2 - ys, x1 - xs:x2 - xs] = res
    m = (bgr2gray_u8(mask_u8) / 255).astype(np.float32)
    body[ys:ye, xs:xe] = blend_linear_u8(face_large, body[ys:ye, xs:xe], m, (1 - m).astype(np.float32))
    return body
