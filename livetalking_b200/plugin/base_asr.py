"""Audio chunk queues and stride context shared by the feature extractors.

Inside LiveTalking this module IS the reference's class: ``from avatars.audio_features.base_asr import BaseASR`` (the
host's own queue / silence / warm-up logic stays untouched, avatars/audio_features/base_asr.py:29-89).  Only when the
plugin is used stand-alone (tests, bench — no LiveTalking checkout on the path) a minimal stand-in with the same
attributes and pacing behaviour is defined:

* 20 ms PCM chunks (320 float32 samples at fps=25) arrive through ``put_audio_frame`` (type 0 = speech);
* ``get_audio_frame`` waits up to 10 ms for a chunk and otherwise hands out a zero chunk of type 1 (silence) — the
  timeout is what paces a silent session;
* ``warm_up`` primes ``stride_left + stride_right`` chunks and drops the first ``stride_left`` from the output queue,
  which delays the audio by the right-context look-ahead.
"""
from __future__ import annotations

import queue
from queue import Queue

import numpy as np

try:
    from avatars.audio_features.base_asr import BaseASR          # the reference's own class (runtime unchanged)
    from avatars.base_avatar import AudioFrameData
    REFERENCE_BASE_ASR = True
except Exception:                                                # stand-alone: no LiveTalking on sys.path
    REFERENCE_BASE_ASR = False
    try:
        from avatars.base_avatar import AudioFrameData           # stubbed base_avatar (tests/stubs.py)
    except Exception:
        from dataclasses import dataclass, field

        @dataclass
        class AudioFrameData:                                    # avatars/base_avatar.py:56-61
            data: np.ndarray
            type: int = 0
            userdata: dict = field(default_factory=dict)

    class BaseASR:
        def __init__(self, opt, parent=None):
            self.opt, self.parent = opt, parent
            self.fps, self.batch_size = opt.fps, opt.batch_size
            self.sample_rate = 16000
            self.chunk = self.sample_rate // (opt.fps * 2)
            self.stride_left_size, self.stride_right_size = opt.l, opt.r
            self.queue, self.output_queue, self.feat_queue = Queue(), Queue(), Queue(maxsize=2)
            self.frames = []

        def flush_talk(self):
            self.queue.queue.clear()

        def put_audio_frame(self, audio_chunk, datainfo: dict):
            self.queue.put(AudioFrameData(data=audio_chunk, type=0, userdata=datainfo))

        def get_audio_frame(self):
            custom = getattr(self.parent, "custom_audiotype", 0) if self.parent else 0
            if custom > 1:
                return AudioFrameData(data=self.parent.get_custom_audio_stream(custom), type=custom, userdata={})
            try:
                return self.queue.get(block=True, timeout=0.01)
            except queue.Empty:
                return AudioFrameData(data=np.zeros(self.chunk, dtype=np.float32), type=1, userdata={})

        def get_audio_out(self):
            return self.output_queue.get()

        def warm_up(self):
            n = self.stride_left_size + self.stride_right_size
            primed = [self.get_audio_frame() for _ in range(n)]
            self.frames.extend(f.data for f in primed)
            for f in primed[self.stride_left_size:]:
                self.output_queue.put(f)

        def run_step(self):
            pass

        def get_next_feat(self, block, timeout):
            return self.feat_queue.get(block, timeout)


def fixed_chunk(data, chunk: int) -> np.ndarray:
    """The reference's ``get_custom_audio_stream`` hands out ``cycle[idx:idx+chunk]`` (avatars/base_avatar.py:303-309), so the
    last chunk of a custom-action clip is SHORT unless the clip length is a multiple of 320.  librosa / the HF feature
    extractor take any length; the engine's feature kernels take a fixed window — pad with zeros (or trim), never raise:
    an exception here would escape ``render()`` and kill the session."""
    a = np.asarray(data, dtype=np.float32).reshape(-1)
    if a.size == chunk:
        return a
    if a.size > chunk:
        return a[:chunk]
    out = np.zeros(chunk, np.float32)
    out[:a.size] = a
    return out
