"""MelASR on the B200 engine — drop-in for avatars/audio_features/mel.py:32-67.

``run_step`` keeps the reference's bookkeeping (2*B chunks pulled and forwarded to ``output_queue``, one list of B
(80,16) windows pushed to ``feat_queue``, the last l+r chunks kept as context) but the feature computation itself —
``audio.melspectrogram`` + window slicing — is one call into the CUDA mel kernels (csrc/mel.cu)."""
from __future__ import annotations

import numpy as np

from .base_asr import BaseASR, fixed_chunk


class MelASR(BaseASR):
    def __init__(self, opt, parent=None, session=None):
        super().__init__(opt, parent)
        self.session = session if session is not None else getattr(parent, "engine_session", None)
        if self.session is None:
            raise RuntimeError("MelASR needs an engine session (no CPU fallback)")

    def run_step(self):
        for _ in range(self.batch_size * 2):
            audioframe = self.get_audio_frame()
            audioframe.data = fixed_chunk(audioframe.data, self.chunk)     # short tail chunk of a custom-action clip
            self.frames.append(audioframe.data)
            self.output_queue.put(audioframe)
        if len(self.frames) <= self.stride_left_size + self.stride_right_size:   # context not enough (mel.py:43-44)
            return
        inputs = np.concatenate(self.frames)
        n_expected = (self.stride_left_size + self.stride_right_size + 2 * self.batch_size) * self.chunk
        if inputs.size != n_expected:
            # run_step before warm_up(): the reference would run librosa on a short buffer; mirror its early return
            # (no features, audio already forwarded) instead of raising inside the render loop
            return
        mel = self.session.mel_step(inputs.astype(np.float32, copy=False))        # (B, 80, 16) float32
        self.feat_queue.put([mel[i] for i in range(self.batch_size)])
        self.frames = self.frames[-(self.stride_left_size + self.stride_right_size):]
