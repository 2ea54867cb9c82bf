"""CPU: weight packer — BN folding and K-major layouts, checked by emulating the engine's gather in numpy
against the oracle on small layers."""
import struct

import numpy as np
import torch
import torch.nn.functional as F

from livetalking_b200 import w2l_pack as WP
from oracle import wav2lip_ref as R


def _blob_entries(blob):
    magic, n, hb = struct.unpack_from("<8sII", blob, 0)
    assert magic[:7] == b"LTBW2L1"
    out = {}
    for i in range(n):
        name, dtype, _, off, nb = struct.unpack_from("<40sIIQQ", blob, 16 + 64 * i)
        out[name.rstrip(b"\0").decode()] = (dtype, off, nb)
    return out


def test_layer_table_agrees_with_oracle():
    a = [(p, k, ci, co, ks) for p, k, ci, co, ks in WP.layer_table()]
    b = [(p, s[0], s[1], s[2], s[3]) for p, s in R.layer_list()]
    assert a == b


def test_blob_layout_and_sizes(w2l_state_dict):
    blob = WP.pack_state_dict(w2l_state_dict)
    ent = _blob_entries(blob)
    assert len(ent) == 54 * 2 + 2
    for i, (prefix, kind, ci, co, k) in enumerate(WP.layer_table()):
        dtype, off, nb = ent[f"L{i:02d}.w"]
        assert off % 256 == 0
        if i == 0:
            assert (dtype, nb) == (1, 32 * 9 * 4)
        elif i == WP.STEM_LAYER:
            assert (dtype, nb) == (0, 16 * 7 * 64 * 2)
        elif i == WP.CONVT4_LAYER:
            assert (dtype, nb) == (0, 16 * 512 * 1024 * 2)
        else:
            assert (dtype, nb) == (0, co * k * k * ci * 2)
    assert ent["head.w"][2] == 96 * 4 and ent["head.b"][2] == 12
    # 53.6 M parameters -> ~107 MB of fp16
    assert 100e6 < len(blob) < 125e6


def test_fold_and_pack_conv_equals_reference_block():
    torch.manual_seed(0)
    ci, co = 8, 16
    sd = {"b.conv_block.0.weight": torch.randn(co, ci, 3, 3), "b.conv_block.0.bias": torch.randn(co),
          "b.conv_block.1.weight": torch.rand(co) + 0.5, "b.conv_block.1.bias": torch.randn(co),
          "b.conv_block.1.running_mean": torch.randn(co), "b.conv_block.1.running_var": torch.rand(co) + 0.5}
    x = torch.randn(2, ci, 6, 5)
    ref = R._block(sd, "b", ("c", ci, co, 3, (1, 1), 1, 0, False), x)
    w, b = WP.fold_bn(sd, "b", "c")
    wp = WP.pack_conv(w)                                           # [co, 9*ci]
    xp = F.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1).numpy().astype(np.float64)   # NHWC
    out = np.zeros((2, 6, 5, co))
    for kh in range(3):
        for kw in range(3):
            t = kh * 3 + kw
            out += xp[:, kh:kh + 6, kw:kw + 5, :] @ wp[:, t * ci:(t + 1) * ci].T
    out = np.maximum(out + b, 0)
    np.testing.assert_allclose(out, ref.permute(0, 2, 3, 1).numpy(), atol=1e-4)


def test_pack_convT_phases_equal_conv_transpose():
    torch.manual_seed(1)
    ci, co, H, W = 8, 4, 3, 5
    w = torch.randn(ci, co, 3, 3)
    x = torch.randn(1, ci, H, W)
    ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1).permute(0, 2, 3, 1).numpy()
    wp = WP.pack_convT_s2(w.numpy().astype(np.float64))            # [co, 9*ci]
    xn = np.zeros((H + 1, W + 1, ci))
    xn[:H, :W] = x[0].permute(1, 2, 0).numpy()
    out = np.zeros((2 * H, 2 * W, co))
    koff = 0
    for a in (0, 1):
        for b in (0, 1):
            for (dy, _) in WP._T_TAPS[a]:
                for (dx, _) in WP._T_TAPS[b]:
                    blk = wp[:, koff:koff + ci]
                    koff += ci
                    out[a::2, b::2] += xn[dy:dy + H, dx:dx + W] @ blk.T
    np.testing.assert_allclose(out, ref[0], atol=1e-5)


def test_pack_stem_and_convT4_layouts():
    w = np.arange(16 * 6 * 7 * 7, dtype=np.float64).reshape(16, 6, 7, 7)
    p = WP.pack_stem(w).reshape(16, 7, 8, 8)
    assert p[3, 2, 5, 4] == w[3, 4, 2, 5] and np.all(p[:, :, 7, :] == 0) and np.all(p[:, :, :, 6:] == 0)
    wt = np.arange(4 * 3 * 4 * 4, dtype=np.float64).reshape(4, 3, 4, 4)   # [Cin, Cout, 4, 4]
    q = WP.pack_convT4(wt)
    assert q.shape == (16 * 3, 4) and q[(2 * 4 + 1) * 3 + 2, 3] == wt[3, 2, 2, 1]
