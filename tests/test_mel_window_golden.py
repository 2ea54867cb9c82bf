"""CPU: the Wav2Lip mel window arithmetic (oracle/mel_ref.py::mel_chunks and csrc/mel.cu::mel_window_kernel) against windows
queued by the reference's OWN MelASR.run_step (avatars/audio_features/mel.py:34-67) — fixture
tests/golden/mel_window_golden.npz from make_golden.py::make_mel_windows (fake mel whose values are the column indices)."""
import os

import numpy as np

from oracle import mel_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mel_window_golden.npz")


def kernel_window_starts(B, l, fps, T):
    """mel.cu:77-85 + launch_mel_step: left = l*80/50, mult = 80/fps (doubles); start = (int)(left + i*mult), tail-clamped."""
    left, mult = max(0.0, l * 80 / 50), 80.0 / fps
    out = []
    for i in range(B):
        s = int(left + i * mult)
        out.append(T - 16 if s + 16 > T else s)
    return out


def test_mel_windows_match_reference_run_step():
    g = np.load(GOLDEN)
    assert len(g["cfgs"]) >= 7
    for B, l, r, fps in g["cfgs"]:
        want = g[f"mel_B{B}_l{l}_r{r}_fps{fps}"]                                    # (B, 16) mel column indices
        n_chunks = int(l + r + 2 * B)
        T = 1 + (n_chunks * 320) // 200
        fake = np.tile(np.arange(T, dtype=np.float64), (80, 1))
        got = np.stack(mel_ref.mel_chunks(fake, n_chunks, int(l), int(r), int(fps)))[:, 0, :].astype(np.int32)
        assert np.array_equal(got, want), (B, l, r, fps)
        assert kernel_window_starts(int(B), int(l), int(fps), T) == want[:, 0].tolist(), (B, l, r, fps)
        assert (np.diff(want, axis=1) == 1).all()                                     # contiguous 16-column windows
    assert g["mel_B2_l4_r4_fps25"][:, 0].tolist() == [4, 4]                           # the tail clamp is exercised
