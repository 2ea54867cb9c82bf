"""GPU: ltb_op_bgr_to_i420 (encoder hand-off) is bit-exact with the OpenCV-pinned oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,H,W", [(1, 2, 4), (3, 48, 64), (2, 90, 36), (1, 720, 1280), (16, 360, 640)])
def test_bgr_to_i420_bit_exact(n, H, W):
    from livetalking_b200 import engine, ops
    from oracle import yuv_ref
    engine.set_device(0)
    ctx = ops.Ctx()
    rng = np.random.default_rng(n * 7 + H)
    frames = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    frames[0, :2, :4] = [[(0, 0, 0), (255, 255, 255), (255, 0, 0), (0, 0, 255)], [(0, 255, 0), (255, 255, 0), (1, 2, 3), (254, 253, 252)]]
    src = ctx.upload(frames)
    dst = ctx.alloc((n, H * 3 // 2, W), np.uint8, zero=True)
    ctx.bgr_to_i420(src, n, H, W, dst)
    got = ctx.download(dst)
    for i in range(n):
        want = yuv_ref.bgr_to_i420(frames[i])
        assert np.array_equal(got[i], want), (i, int(np.abs(got[i].astype(int) - want).max()))
    ctx.close()


def test_bgr_to_i420_rejects_odd_sizes():
    from livetalking_b200 import engine, ops
    engine.set_device(0)
    ctx = ops.Ctx()
    src = ctx.alloc((1, 6, 6, 3), np.uint8, zero=True)
    dst = ctx.alloc((1, 9, 6), np.uint8, zero=True)
    with pytest.raises(engine.LtbError):
        ctx.bgr_to_i420(src, 1, 6, 6, dst)        # width not a multiple of 4
    with pytest.raises(engine.LtbError):
        ctx.bgr_to_i420(src, 1, 5, 8, dst)        # odd height
    ctx.close()
