"""CPU: oracle/mel_ref.py::melspectrogram against the reference's OWN audio.melspectrogram (avatars/wav2lip/audio.py +
hparams.py executed here; only the absent librosa's stft / filters.mel were substituted by the oracle's restatements,
whose arguments the fixture generator asserts) — tests/golden/mel_chain_golden.npz, make_golden.py::make_mel_chain."""
import os

import numpy as np

from oracle import mel_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mel_chain_golden.npz")


def test_melspectrogram_matches_reference_audio_py():
    g = np.load(GOLDEN)
    mel = mel_ref.melspectrogram(g["pcm"])
    want = g["mel"]
    assert mel.shape == want.shape == (80, 84)
    assert np.abs(mel - want).max() <= 1e-9          # same float64 chain: pre-emphasis, |STFT|, mel, dB, -ref, normalise, clip
    assert (want == -4.0).mean() > 0.05 and want.max() > 2.5   # lower clip and the loud end of the range are exercised
