"""CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py): Wav2Lip mel front-end.

float64 numpy restatement of

    avatars/wav2lip/audio.py:20-23   preemphasis  = scipy.signal.lfilter([1,-k],[1],wav)
    avatars/wav2lip/audio.py:45-51   melspectrogram
    avatars/wav2lip/audio.py:57-61   _stft -> librosa.stft(n_fft=800, hop=200, win=800)
    avatars/wav2lip/audio.py:92-105  _linear_to_mel, _build_mel_basis, _amp_to_db
    avatars/wav2lip/audio.py:110-114 _normalize (symmetric, clipping)
    avatars/wav2lip/hparams.py:33-73 constants
    avatars/audio_features/mel.py:47-63  window slicing of MelASR.run_step

librosa (third-party, unpinned in requirements.txt:44, NOT installed here) provides
``stft`` and ``filters.mel``; their published algorithms are restated below
(centre-padded periodic-Hann STFT; Slaney mel scale with Slaney area normalisation).
**parity unpinned** against a real librosa for these two functions; cross-checked in tests
against ``torch.stft`` and ``torchaudio.functional.melscale_fbanks``.  Everything around them
(pre-emphasis, dB, normalisation, clipping, the call arguments, the window slicing) is pinned to the
reference's own audio.py / hparams.py / MelASR.run_step executed in the build container
(tests/golden/mel_chain_golden.npz, mel_window_golden.npz).  The frames touched by
the centre padding (0,1,T-2,T-1) are never selected by the window slicing, so librosa's
``pad_mode`` default (constant vs reflect across versions) does not reach the output.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 800
HOP = 200
WIN = 800
N_MELS = 80
FMIN = 55.0
FMAX = 7600.0
PREEMPH = 0.97
MIN_LEVEL_DB = -100.0
REF_LEVEL_DB = 20.0
MAX_ABS = 4.0
MEL_STEP = 16


def preemphasis(wav: np.ndarray) -> np.ndarray:
    """audio.py:20-23 — y[n] = x[n] - 0.97 x[n-1], float64 (lfilter promotes)."""
    x = np.asarray(wav, dtype=np.float64)
    y = x.copy()
    y[1:] -= PREEMPH * x[:-1]
    return y


def hann_periodic(n: int = WIN) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True) = librosa's default window."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft_mag(y: np.ndarray, pad_mode: str = "constant") -> np.ndarray:
    """|librosa.stft(y, n_fft=800, hop_length=200, win_length=800, center=True)| -> (401, T)."""
    y = np.asarray(y, dtype=np.float64)
    pad = N_FFT // 2
    if pad_mode == "constant":
        yp = np.concatenate([np.zeros(pad), y, np.zeros(pad)])
    else:
        yp = np.pad(y, pad, mode=pad_mode)
    T = 1 + len(y) // HOP
    win = hann_periodic()
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(T)[:, None]
    frames = yp[idx] * win[None, :]
    return np.abs(np.fft.rfft(frames, n=N_FFT, axis=1)).T


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-12) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis() -> np.ndarray:
    """librosa.filters.mel(sr=16000, n_fft=800, n_mels=80, fmin=55, fmax=7600) -> (80,401) float32
    (htk=False Slaney scale, norm='slaney'); audio.py:98-101."""
    fftfreqs = np.linspace(0.0, SAMPLE_RATE / 2.0, 1 + N_FFT // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(FMIN), _hz_to_mel(FMAX), N_MELS + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((N_MELS, 1 + N_FFT // 2))
    for i in range(N_MELS):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:N_MELS + 2] - mel_f[:N_MELS])
    w *= enorm[:, None]
    return w.astype(np.float32)   # librosa returns float32


def melspectrogram(wav: np.ndarray) -> np.ndarray:
    """audio.py:45-51 -> (80, 1 + len//200) float64 in [-4, 4]."""
    D = stft_mag(preemphasis(wav))
    S = np.dot(mel_basis(), D)                                           # float32 @ float64 -> float64
    min_level = np.exp(MIN_LEVEL_DB / 20.0 * np.log(10.0))               # audio.py:104
    S = 20.0 * np.log10(np.maximum(min_level, S)) - REF_LEVEL_DB        # audio.py:105,47
    return np.clip((2 * MAX_ABS) * ((S - MIN_LEVEL_DB) / (-MIN_LEVEL_DB)) - MAX_ABS, -MAX_ABS, MAX_ABS)


def mel_chunks(mel: np.ndarray, n_frames_total: int, stride_left: int, stride_right: int, fps: int = 25):
    """mel.py:47-63: slice windows of 16 mel columns, one per video frame.

    n_frames_total = len(self.frames) (20 ms chunks in the buffer)."""
    T = mel.shape[1]
    left = max(0, stride_left * 80 / 50)
    mult = 80.0 / fps
    chunks = []
    i = 0
    while i < (n_frames_total - stride_left - stride_right) / 2:
        s = int(left + i * mult)
        if s + MEL_STEP > T:
            chunks.append(mel[:, T - MEL_STEP:])
        else:
            chunks.append(mel[:, s:s + MEL_STEP])
        i += 1
    return chunks


def mel_step(pcm: np.ndarray, batch: int, stride_left: int = 10, stride_right: int = 10, fps: int = 25) -> np.ndarray:
    """One MelASR.run_step feature computation on a full buffer of (l + r + 2B) chunks of
    320 samples -> (B, 80, 16) float64."""
    n_chunks = len(pcm) // 320
    mel = melspectrogram(pcm)
    ch = mel_chunks(mel, n_chunks, stride_left, stride_right, fps)
    assert len(ch) == batch, (len(ch), batch)
    return np.stack(ch, 0)
