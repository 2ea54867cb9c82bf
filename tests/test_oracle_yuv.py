"""CPU: the I420 oracle (oracle/yuv_ref.py) is pinned bit-exactly against the installed OpenCV (COLOR_BGR2YUV_I420)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import yuv_ref  # noqa: E402


@pytest.mark.parametrize("shape", [(2, 2), (2, 4), (6, 10), (48, 64), (90, 34), (360, 640), (720, 1280)])
def test_i420_oracle_matches_opencv(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
    assert np.array_equal(yuv_ref.bgr_to_i420(img), cv2.cvtColor(img, cv2.COLOR_BGR2YUV_I420))


def test_i420_oracle_saturation_and_primaries():
    cols = [(0, 0, 0), (255, 255, 255), (255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 255, 0), (0, 255, 255), (255, 0, 255), (1, 254, 3)]
    for c in cols:
        img = np.empty((4, 8, 3), np.uint8)
        img[:] = c
        img[1, 3] = (255 - c[0], c[1], 255 - c[2])          # a pixel that is NOT a chroma sample: must only move Y
        got = yuv_ref.bgr_to_i420(img)
        assert np.array_equal(got, cv2.cvtColor(img, cv2.COLOR_BGR2YUV_I420)), c
    with pytest.raises(AssertionError):
        yuv_ref.bgr_to_i420(np.zeros((3, 4, 3), np.uint8))    # odd height
