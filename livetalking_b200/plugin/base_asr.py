"""Audio chunk queues and stride context shared by the feature extractors — host logic, same names, argument meaning
and pacing behaviour as the reference's ``BaseASR`` (avatars/audio_features/base_asr.py:29-89):

* 20 ms PCM chunks (320 float32 samples at fps=25) arrive through ``put_audio_frame`` (type 0 = speech);
* ``get_audio_frame`` waits up to 10 ms for a chunk and otherwise synthesises a zero chunk of type 1 (silence) —
  this timeout is what paces a silent session (base_asr.py:57-69);
* ``warm_up`` primes ``stride_left + stride_right`` chunks and drops the first ``stride_left`` from the output queue,
  which delays the audio by the right-context look-ahead (base_asr.py:76-82).
"""
from __future__ import annotations

import queue
from dataclasses import dataclass, field
from queue import Queue

import numpy as np

try:  # inside LiveTalking the reference's dataclass is used so isinstance checks keep working
    from avatars.base_avatar import AudioFrameData  # type: ignore
except Exception:  # stand-alone (tests, bench): identical fields (avatars/base_avatar.py:56-61)
    @dataclass
    class AudioFrameData:  # type: ignore[no-redef]
        data: np.ndarray
        type: int = 0
        userdata: dict = field(default_factory=dict)


class BaseASR:
    def __init__(self, opt, parent=None):
        self.opt = opt
        self.parent = parent
        self.fps = opt.fps
        self.sample_rate = 16000
        self.chunk = self.sample_rate // (opt.fps * 2)      # 320 samples = 20 ms
        self.queue: Queue = Queue()
        self.output_queue: Queue = Queue()
        self.batch_size = opt.batch_size
        self.frames = []
        self.stride_left_size = opt.l
        self.stride_right_size = opt.r
        self.feat_queue: Queue = Queue(maxsize=2)

    def flush_talk(self):
        self.queue.queue.clear()

    def put_audio_frame(self, audio_chunk, datainfo: dict):
        self.queue.put(AudioFrameData(data=audio_chunk, type=0, userdata=datainfo))

    def get_audio_frame(self) -> AudioFrameData:
        try:
            if self.parent and getattr(self.parent, "custom_audiotype", 0) > 1:
                frame = self.parent.get_custom_audio_stream(self.parent.custom_audiotype)
                return AudioFrameData(data=frame, type=self.parent.custom_audiotype, userdata={})
            return self.queue.get(block=True, timeout=0.01)
        except queue.Empty:
            return AudioFrameData(data=np.zeros(self.chunk, dtype=np.float32), type=1, userdata={})

    def get_audio_out(self) -> AudioFrameData:
        return self.output_queue.get()

    def warm_up(self):
        for _ in range(self.stride_left_size + self.stride_right_size):
            audio_frame = self.get_audio_frame()
            self.frames.append(audio_frame.data)
            self.output_queue.put(audio_frame)
        for _ in range(self.stride_left_size):
            self.output_queue.get()

    def run_step(self):
        pass

    def get_next_feat(self, block, timeout):
        return self.feat_queue.get(block, timeout)
