"""CPU: the oracle's wav2lip pipeline glue (batch assembly a4 + forward a5 + paste-back a6) against outputs of the reference's
OWN LipReal.inference_batch / paste_back_frame (avatars/wav2lip_avatar.py:116-147), captured by
tests/golden/make_golden.py::make_lipreal in tests/golden/lipreal_golden.npz."""
import os

import numpy as np
import torch

from oracle import paste_ref as P
from oracle import wav2lip_ref as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lipreal_golden.npz")


def test_oracle_pipeline_equals_reference_lipreal(w2l_state_dict):
    g = np.load(GOLDEN)
    seed, index, B = int(g["seed"]), int(g["index"]), int(g["batch"])
    n = 2
    mel, img = R.synth_inputs(n, seed=seed)
    faces = list((img[:, 3:6].permute(0, 2, 3, 1).numpy() * 255.0).round().astype(np.uint8))
    rng = np.random.default_rng(seed)
    frames = [rng.integers(0, 256, (120, 160, 3), dtype=np.uint8) for _ in range(n)]
    coords = [tuple(int(v) for v in c) for c in g["coords"]]
    melB = np.tile(mel.numpy().reshape(n, 80, 16), (2, 1, 1))[:B]
    batch = P.w2l_build_batch(faces, index, B)                                          # a4
    out = R.wav2lip_forward(w2l_state_dict, torch.from_numpy(melB).reshape(B, 1, 80, 16), torch.from_numpy(batch))   # a5
    pred = out.numpy().transpose(0, 2, 3, 1) * 255.0
    # CPU conv algorithms / blockings differ with batch size and thread count (1 thread vs many: up to ~0.25 of 255 after
    # 54 layers of fp32), so values are compared with tolerances; the u8 conversion may then flip by one LSB at integer crossings
    assert np.abs(pred[:, ::8, ::8, :] - g["pred_sub"]).max() <= 0.75
    assert np.abs(pred[:, ::8, ::8, :] - g["pred_sub"]).mean() <= 0.02
    d8 = np.abs(pred.astype(np.uint8)[:, ::4, ::4, :].astype(int) - g["pred_u8_sub"].astype(int))
    assert d8.max() <= 1 and (d8 > 0).mean() < 0.05
    for i in range(B):                                                                   # a6
        idx = P.mirror_index(n, index + i)
        y1, y2, x1, x2 = coords[idx]
        got = P.w2l_paste_back(pred[i], frames[idx], coords[idx])
        outside = got.copy()
        outside[y1:y2, x1:x2] = frames[idx][y1:y2, x1:x2]
        assert np.array_equal(outside, frames[idx])                                      # untouched outside the bbox, fresh copy
        dc = np.abs(got[y1:y2, x1:x2].astype(int) - g[f"crop{i}"].astype(int))
        assert dc.shape == (y2 - y1, x2 - x1, 3) and dc.max() <= 1 and (dc > 0).mean() < 0.08, (i, dc.max(), (dc > 0).mean())
