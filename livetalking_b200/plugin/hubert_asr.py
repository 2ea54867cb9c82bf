"""HubertASR on the B200 engine — drop-in for avatars/audio_features/hubert.py:13-51.

Same bookkeeping as the reference's ``run_step``: 2*B chunks forwarded to ``output_queue``, ``is_all_silence`` tracked over the
batch, one list of B (16, 1024) float32 windows queued — the zero windows ``B*[np.zeros((10,1024))]`` when this batch AND the
previous one were silent (hubert.py:40-41: the reference then skips HuBERT; the zero default has 10 rows there and is never
consumed with speech) — and l+r chunks of context kept.  ``get_hubert_from_16k_speech`` + ``_feature2chunks`` are one CUDA-graph
launch (livetalking_b200/hubert.py)."""
from __future__ import annotations

import numpy as np

from .base_asr import BaseASR, fixed_chunk


class HubertASR(BaseASR):
    def __init__(self, opt, parent, audio_processor, audio_feat_length=(4, 4)):
        super().__init__(opt, parent)
        self.audio_processor = audio_processor          # livetalking_b200.hubert.HubertFeatures
        if audio_processor is None:
            raise RuntimeError("HubertASR needs an engine HubertFeatures object (no CPU fallback)")
        if tuple(audio_feat_length) != (4, 4):
            raise ValueError("the engine extractor is built for audio_feat_length [4,4] (LightReal, ultralight_avatar.py:137)")
        self.audio_feat_length = list(audio_feat_length)
        self.last_is_silence = True

    def run_step(self):
        is_all_silence = True
        for _ in range(self.batch_size * 2):
            audio_frame = self.get_audio_frame()
            if audio_frame.type == 0:
                is_all_silence = False
            audio_frame.data = fixed_chunk(audio_frame.data, self.chunk)    # short tail chunk of a custom-action clip
            self.frames.append(audio_frame.data)
            self.output_queue.put(audio_frame)
        if len(self.frames) <= self.stride_left_size + self.stride_right_size:
            return
        mel_chunks = self.batch_size * [np.zeros((10, 1024), dtype=np.float32)]       # hubert.py:40 (silence default)
        if not is_all_silence or not self.last_is_silence:
            inputs = np.concatenate(self.frames)
            n_expected = (self.stride_left_size + self.stride_right_size + 2 * self.batch_size) * self.chunk
            if inputs.size != n_expected:              # run_step before warm_up(): no features yet, as the early return above
                return
            feats = self.audio_processor.run(inputs.astype(np.float32, copy=False))   # (B, 16, 1024) float32
            mel_chunks = [feats[i] for i in range(self.batch_size)]
        self.feat_queue.put(mel_chunks)
        self.frames = self.frames[-(self.stride_left_size + self.stride_right_size):]
        self.last_is_silence = is_all_silence
