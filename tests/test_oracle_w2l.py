"""CPU: the wav2lip256 oracle against the golden vectors produced by the UNMODIFIED reference module
(tests/golden/make_golden.py), plus invariants of the synthetic weights."""
import os

import numpy as np
import torch

from oracle import wav2lip_ref as R


def test_layer_table_matches_reference_counts():
    # 54 conv+BN blocks, 27.789 GMAC per frame, 380 state_dict tensors
    layers = R.layer_list()
    assert len(layers) == 54
    assert len(R._state_shapes()) == 54 * 7 + 2 == 380


def test_oracle_matches_reference_golden(w2l_state_dict, golden_dir):
    g = np.load(os.path.join(golden_dir, "w2l_golden.npz"))
    mel, img = R.synth_inputs(1, seed=int(g["seed"][1]))
    taps = {}
    out = R.wav2lip_forward(w2l_state_dict, mel, img, taps)
    pred = out.numpy().transpose(0, 2, 3, 1) * 255.0
    # same torch build / same CPU kernels => essentially bit-identical; allow 1 LSB flips from summation order
    diff = np.abs(pred.astype(np.uint8).astype(int) - g["pred_u8"].astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
    np.testing.assert_allclose(pred[:, ::4, ::4, :], g["pred_f32_sub"], atol=2e-2)
    names = [p for p, _ in R.layer_list()]
    stats = np.array([[float(taps[n].mean()), float(taps[n].abs().mean()), float(taps[n].std())] for n in names])
    np.testing.assert_allclose(stats, g["layer_stats"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(taps["audio_encoder.12"].numpy().reshape(-1), g["audio_emb"], rtol=1e-3, atol=1e-4)


def test_synthetic_weights_are_conditioned(w2l_state_dict):
    mel, img = R.synth_inputs(1, seed=3)
    taps = {}
    out = R.wav2lip_forward(w2l_state_dict, mel, img, taps)
    assert 0.2 < float(out.std()) < 0.45 and float(out.min()) < 0.02 and float(out.max()) > 0.98
    for k, v in taps.items():
        assert float(v.abs().max()) < 64.0, k          # fp16-safe activations
    # the output must depend on the audio input (otherwise parity on the audio branch is vacuous)
    mel2 = mel.roll(3, dims=3) * -1.0
    out2 = R.wav2lip_forward(w2l_state_dict, mel2, img)
    assert float((out - out2).abs().mean()) > 1e-3


def test_psnr_helper():
    a = np.zeros((4, 4), np.uint8)
    b = a.copy()
    b[0, 0] = 16
    assert abs(R.psnr_u8(a, b) - 10 * np.log10(255 ** 2 / 16.0)) < 1e-9
    assert R.psnr_u8(a, a) == float("inf")
