"""CPU: oracle.musetalk_ref.positional_encoding against the reference's OWN PositionalEncoding module
(avatars/musetalk/models/unet.py:12-27; fixture tests/golden/pe_golden.npz from make_golden.py::make_pe)."""
import os

import numpy as np
import torch

from oracle import musetalk_ref as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pe_golden.npz")


def test_positional_encoding_matches_reference_module():
    g = np.load(GOLDEN)
    x = torch.randn(2, 50, 384, generator=torch.Generator().manual_seed(int(g["seed"])))
    y = M.positional_encoding(x).numpy()
    assert np.array_equal(y[:, ::5, ::7], g["y_sub"])                      # same fp32 arithmetic: bit-equal
    assert np.array_equal((y - x.numpy())[0], (torch.from_numpy(g["table"]) + x[0]).numpy() - x[0].numpy())
    assert np.abs(M.positional_encoding(torch.zeros(1, 50, 384))[0].numpy() - g["table"]).max() == 0.0
