"""CPU: mel oracle vs an independent torch/torchaudio pipeline (golden) and vs scipy / torch primitives."""
import os

import numpy as np
import pytest

from oracle import mel_ref as M


def test_mel_matches_independent_pipeline_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "mel_golden.npz"))
    win = M.mel_step(g["pcm"], 16)
    assert win.shape == (16, 80, 16)
    np.testing.assert_allclose(win, g["windows"], atol=2e-5)
    np.testing.assert_allclose(M.melspectrogram(g["pcm"]), g["mel"], atol=2e-5)


def test_primitives_against_scipy_and_torch():
    scipy_signal = pytest.importorskip("scipy.signal")
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(1)
    x = rng.standard_normal(7040).astype(np.float32) * 0.1
    y = M.preemphasis(x)
    assert y.dtype == np.float64
    assert np.array_equal(y, scipy_signal.lfilter([1, -0.97], [1], x))
    D = M.stft_mag(y)
    assert D.shape == (401, 36)
    Dt = torch.stft(torch.from_numpy(y), 800, 200, 800, window=torch.hann_window(800, periodic=True, dtype=torch.float64),
                    center=True, pad_mode="constant", return_complex=True).abs().numpy()
    np.testing.assert_allclose(D, Dt, atol=1e-10)
    fb = M.mel_basis()
    assert fb.shape == (80, 401) and fb.dtype == np.float32
    try:
        import torchaudio
        ref = torchaudio.functional.melscale_fbanks(401, 55.0, 7600.0, 80, 16000, norm="slaney", mel_scale="slaney").T.numpy()
        np.testing.assert_allclose(fb, ref, atol=2e-7)
    except ImportError:
        pass


def test_window_slicing_rules():
    # mel.py:50-63 with l=r=10, fps=25: starts int(16+3.2 i); tail clamps to the last 16 columns
    mel = np.arange(80 * 84, dtype=np.float64).reshape(80, 84)
    ch = M.mel_chunks(mel, 20 + 32, 10, 10, 25)
    assert len(ch) == 16
    assert [int(c[0, 0]) for c in ch] == [int(16 + 3.2 * i) for i in range(16)]
    short = mel[:, :30]
    ch = M.mel_chunks(short, 20 + 8, 10, 10, 25)
    assert [int(c[0, 0]) for c in ch] == [14, 14, 14, 14]   # start+16 > 30 -> last 16 columns
    # silence (all zeros) sits on the clip floor
    z = M.mel_step(np.zeros(7040, np.float32), 1)
    assert z.shape == (1, 80, 16) and np.all(z == -4.0)
