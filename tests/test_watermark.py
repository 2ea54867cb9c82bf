"""The device watermark: the pixel list rasterised once by OpenCV reproduces cv2.putText on any frame (CPU), and
the device stamp writes exactly those pixels (GPU)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from livetalking_b200 import watermark as WM  # noqa: E402


@pytest.mark.parametrize("shape", [(720, 1280), (64, 96), (18, 40)], ids=["720p", "small", "clipped"])
def test_pixel_list_reproduces_cv2_puttext(shape):
    H, W = shape
    rng = np.random.default_rng(H)
    frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    want = frame.copy()
    cv2.putText(want, "LiveTalking", (10, 20), cv2.FONT_HERSHEY_SIMPLEX, 0.3, (128, 128, 128), 1)      # avatars/base_avatar.py:449
    pix = WM.text_pixels(H, W)
    assert pix.dtype == np.int32 and pix.ndim == 2 and pix.shape[1] == 2 and len(pix) > 20
    assert np.array_equal(WM.stamp_host(frame.copy(), pix), want)


@pytest.mark.gpu
def test_device_stamp_matches_cv2_then_i420():
    from livetalking_b200 import engine
    from livetalking_b200.ops import Ctx
    from oracle import yuv_ref
    engine.set_device(0)
    N, H, W = 3, 72, 128
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
    want = frames.copy()
    for f in want:
        cv2.putText(f, "LiveTalking", (10, 20), cv2.FONT_HERSHEY_SIMPLEX, 0.3, (128, 128, 128), 1)
    ctx = Ctx()
    d = ctx.upload(frames)
    pix = ctx.upload(WM.text_pixels(H, W))
    ctx.stamp_pixels(d, N, H, W, pix)
    assert np.array_equal(ctx.download(d), want)
    yuv = ctx.alloc((N, H * 3 // 2, W), np.uint8, zero=True)
    ctx.bgr_to_i420(d, N, H, W, yuv)
    got = ctx.download(yuv)
    for i in range(N):
        assert np.array_equal(got[i], yuv_ref.bgr_to_i420(want[i]))
    ctx.close()
