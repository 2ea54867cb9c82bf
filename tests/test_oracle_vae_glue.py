"""CPU: the MuseTalk VAE pre/post-processing of the oracle (oracle/musetalk_ref.py::preprocess_img, decode_latents_u8)
against outputs of the reference's OWN wrapper methods (avatars/musetalk/models/vae.py:51-82, :96-108), captured in
tests/golden/vae_glue_golden.npz by make_golden.py::make_vae_glue (fake decoder: the diffusers network itself is absent)."""
import os
import zlib

import numpy as np
import torch

from oracle import musetalk_ref as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vae_glue_golden.npz")


def test_preprocess_img_matches_reference_wrapper():
    g = np.load(GOLDEN)
    img = np.random.default_rng(int(g["seed"])).integers(0, 256, (256, 256, 3), dtype=np.uint8)
    for half, key in ((False, "pre_full"), (True, "pre_half")):
        x = M.preprocess_img(img, half).numpy()
        assert x.shape == (1, 3, 256, 256) and x.dtype == np.float32
        assert np.array_equal(x[0, :, ::37, ::41], g[key + "_sub"])
        assert zlib.crc32(np.ascontiguousarray(x).tobytes()) == int(g[key + "_crc"])      # every float bit-equal
    assert (M.preprocess_img(img, True)[0, :, 128:] == -1.0).all()                           # lower half masked to 0 -> -1 after Normalize


def test_decode_post_processing_matches_reference_wrapper(monkeypatch):
    g = np.load(GOLDEN)
    sample = torch.randn(2, 3, 64, 48, generator=torch.Generator().manual_seed(int(g["seed"]))) * 0.8
    monkeypatch.setattr(M, "vae_decode", lambda sd, cfg, z: sample)                          # same fake decoder as the fixture
    out = M.decode_latents_u8(None, M.VAE_SMALL, torch.zeros(2, 4, 8, 6))
    assert out.dtype == np.uint8 and out.shape == (2, 64, 48, 3)
    assert np.array_equal(out, g["post"])                                                    # clamp, x255, round-half-even, RGB->BGR
    assert out.min() == 0 and out.max() == 255                                               # the clamp is exercised
