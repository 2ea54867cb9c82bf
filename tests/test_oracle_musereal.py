"""CPU: what the reference's OWN MuseReal.inference_batch (avatars/musetalk_avatar.py:130-152) feeds the UNet — fixture
tests/golden/musereal_golden.npz (make_golden.py::make_musereal, UNet / VAE replaced by recorders) — against the glue the
oracle-side tests and the engine use: latents gathered by mirror_index, audio features + positional encoding, timestep 0."""
import os

import numpy as np
import torch

from oracle import musetalk_ref as M
from oracle.paste_ref import mirror_index

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "musereal_golden.npz")


def test_unet_inputs_match_reference_inference_batch():
    g = np.load(GOLDEN)
    seed, index, B, n = int(g["seed"]), int(g["index"]), int(g["batch"]), int(g["n"])
    gen = torch.Generator().manual_seed(seed)
    lat_list = [torch.randn(1, 8, 4, 4, generator=gen) for _ in range(n)]
    aud = torch.stack([torch.randn(50, 384, generator=gen) for _ in range(B)])
    idxs = [mirror_index(n, index + i) for i in range(B)]                      # 4 -> [1, 0, 0, 1, 2] over a 3-frame avatar
    assert idxs == [1, 0, 0, 1, 2]
    lat = torch.cat([lat_list[i] for i in idxs], 0)
    assert np.array_equal(lat.numpy(), g["latents"])
    ctx = M.positional_encoding(aud)
    assert np.array_equal(ctx.numpy()[:, ::7, ::11], g["ctx_sub"])
    assert g["timesteps"].tolist() == [0]                                      # constant timestep (musetalk_avatar.py:61)
    # the recorder UNet returned latents[:, :4] * 2 + 1 and the fake VAE passed it through: the wrapper adds nothing else
    assert np.array_equal(g["out"], (lat[:, :4] * 2.0 + 1.0).numpy())
