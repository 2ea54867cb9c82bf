"""CPU: host-side logic of the plugin layer (queues, silence synthesis, warm-up, run_step bookkeeping), and — when the
reference tree is present (build container) — equivalence with the reference's own BaseASR on the same event sequence."""
import importlib.util
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import stubs  # noqa: E402

stubs.install()
from livetalking_b200.plugin import base_asr as B  # noqa: E402


def _feed(asr, n, rng):
    chunks = [rng.standard_normal(320).astype(np.float32) for _ in range(n)]
    for i, c in enumerate(chunks):
        asr.put_audio_frame(c, {"i": i})
    return chunks


def test_silence_synthesis_and_warm_up():
    opt = stubs.Opt(batch_size=2)
    a = B.BaseASR(opt)
    assert a.chunk == 320 and a.feat_queue.maxsize == 2
    f = a.get_audio_frame()                             # empty queue -> zeros, type 1 (base_asr.py:66-69)
    assert f.type == 1 and f.data.shape == (320,) and f.data.dtype == np.float32 and not f.data.any()
    rng = np.random.default_rng(0)
    chunks = _feed(a, 25, rng)
    a.warm_up()                                         # l + r = 20 chunks consumed, first l = 10 dropped from output
    assert len(a.frames) == 20 and a.output_queue.qsize() == 10
    out = a.get_audio_out()
    assert out.type == 0 and np.array_equal(out.data, chunks[10]) and out.userdata == {"i": 10}
    a.flush_talk()
    assert a.queue.qsize() == 0


def test_custom_audio_stream_has_priority():
    class Parent:
        custom_audiotype = 2

        def get_custom_audio_stream(self, t):
            return np.full(320, 0.25, np.float32)

    a = B.BaseASR(stubs.Opt(), Parent())
    a.put_audio_frame(np.ones(320, np.float32), {})
    f = a.get_audio_frame()
    assert f.type == 2 and float(f.data[0]) == 0.25      # base_asr.py:59-62


@pytest.mark.skipif(not os.path.exists("/root/reference/avatars/audio_features/base_asr.py"), reason="reference tree not present")
def test_same_behaviour_as_reference_base_asr():
    spec = importlib.util.spec_from_file_location("ref_base_asr", "/root/reference/avatars/audio_features/base_asr.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)                         # imports the stubbed avatars.base_avatar
    opt = stubs.Opt(batch_size=3)
    ours, theirs = B.BaseASR(opt), ref.BaseASR(opt)
    rng1, rng2 = np.random.default_rng(5), np.random.default_rng(5)
    _feed(ours, 23, rng1)
    _feed(theirs, 23, rng2)
    ours.warm_up()
    theirs.warm_up()
    for _ in range(8):                                   # drains speech then synthesises silence
        a, b = ours.get_audio_frame(), theirs.get_audio_frame()
        assert a.type == b.type and np.array_equal(a.data, b.data) and a.userdata == b.userdata
    assert ours.output_queue.qsize() == theirs.output_queue.qsize()
    assert len(ours.frames) == len(theirs.frames) and all(np.array_equal(x, y) for x, y in zip(ours.frames, theirs.frames))


def test_mel_asr_run_step_bookkeeping_with_fake_session():
    """run_step: 2B chunks forwarded, one feature list queued, l+r chunks of context kept (mel.py:36-67, Appendix E)."""
    from livetalking_b200.plugin.mel_asr import MelASR

    class FakeSession:
        def __init__(self):
            self.calls = []

        def mel_step(self, pcm):
            self.calls.append(pcm.copy())
            return np.zeros((2, 80, 16), np.float32)

    opt = stubs.Opt(batch_size=2)
    sess = FakeSession()
    asr = MelASR(opt, None, sess)
    rng = np.random.default_rng(1)
    chunks = _feed(asr, 24, rng)
    asr.warm_up()
    asr.run_step()
    assert asr.feat_queue.qsize() == 1 and asr.output_queue.qsize() == 10 + 4
    feats = asr.feat_queue.get()
    assert len(feats) == 2 and feats[0].shape == (80, 16)
    assert sess.calls[0].size == (10 + 10 + 4) * 320 and np.array_equal(sess.calls[0], np.concatenate(chunks[:24]))
    assert len(asr.frames) == 20 and np.array_equal(asr.frames[0], chunks[4])
    with pytest.raises(RuntimeError):
        MelASR(opt, None, None)                          # no engine session -> loud failure, never a CPU fallback


def test_whisper_asr_run_step_bookkeeping_matches_reference():
    """WhisperASR.run_step (whisper.py:58-76): 2B chunks forwarded, the whole l+r+2B context handed to the feature extractor,
    one list of B (50, 384) arrays queued, l+r chunks kept.  In the build container the reference's own class runs the same
    event sequence (its Audio2Feature replaced by a recorder) and every queue / buffer must match."""
    from livetalking_b200.plugin.whisper_asr import WhisperASR

    class FakeFeatures:                                   # stands in for livetalking_b200.whisper.WhisperFeatures
        def __init__(self, B):
            self.B, self.calls = B, []

        def run(self, pcm):
            self.calls.append(pcm.copy())
            return np.zeros((self.B, 50, 384), np.float16)

    B = 3
    opt = stubs.Opt(batch_size=B)
    fake = FakeFeatures(B)
    ours = WhisperASR(opt, None, fake)
    rng = np.random.default_rng(2)
    chunks = _feed(ours, 20 + 2 * B + 2, rng)
    ours.warm_up()
    ours.run_step()
    assert ours.feat_queue.qsize() == 1 and ours.output_queue.qsize() == 10 + 2 * B
    feats = ours.feat_queue.get()
    assert len(feats) == B and feats[0].shape == (50, 384)
    assert fake.calls[0].dtype == np.float32 and np.array_equal(fake.calls[0], np.concatenate(chunks[:20 + 2 * B]))
    assert len(ours.frames) == 20 and np.array_equal(ours.frames[0], chunks[2 * B])
    with pytest.raises(RuntimeError):
        WhisperASR(opt, None, None)                       # no engine object -> loud failure, never a CPU fallback

    ref_path = "/root/reference/avatars/audio_features/whisper.py"
    if not os.path.exists(ref_path):
        return
    import types
    a2f = types.ModuleType("avatars.musetalk.whisper.audio2feature")
    a2f.Audio2Feature = object
    for name in ("avatars.musetalk", "avatars.musetalk.whisper"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["avatars.musetalk.whisper.audio2feature"] = a2f
    spec = importlib.util.spec_from_file_location("ref_base_asr2", "/root/reference/avatars/audio_features/base_asr.py")
    ref_base = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_base)
    af = types.ModuleType("avatars.audio_features")
    af.__path__ = []
    sys.modules["avatars.audio_features"] = af
    sys.modules["avatars.audio_features.base_asr"] = ref_base
    spec = importlib.util.spec_from_file_location("ref_whisper_asr", ref_path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    class Recorder:                                        # the reference's audio_processor: (T, 5, 384) hidden-state stack
        def __init__(self):
            self.calls = []

        def audio2feat(self, pcm):
            self.calls.append(np.asarray(pcm).copy())
            return np.zeros((1500, 5, 384), np.float32)

    rec = Recorder()
    theirs = ref.WhisperASR(opt, None, rec)
    _feed(theirs, 20 + 2 * B + 2, np.random.default_rng(2))
    theirs.warm_up()
    theirs.run_step()
    fake2 = FakeFeatures(B)
    again = WhisperASR(opt, None, fake2)
    _feed(again, 20 + 2 * B + 2, np.random.default_rng(2))
    again.warm_up()
    again.run_step()
    assert theirs.feat_queue.qsize() == again.feat_queue.qsize() == 1
    assert theirs.output_queue.qsize() == again.output_queue.qsize()
    tf, of = theirs.feat_queue.get(), again.feat_queue.get()
    assert len(tf) == len(of) == B and tf[0].shape == of[0].shape == (50, 384)
    assert np.array_equal(rec.calls[0], fake2.calls[0])                                # identical PCM context handed over
    assert len(theirs.frames) == len(again.frames) and all(np.array_equal(x, y) for x, y in zip(theirs.frames, again.frames))
    for _ in range(theirs.output_queue.qsize()):
        a, b = theirs.output_queue.get(), again.output_queue.get()
        assert a.type == b.type and np.array_equal(a.data, b.data)
    for k in ("avatars.audio_features", "avatars.audio_features.base_asr", "avatars.musetalk.whisper.audio2feature"):
        sys.modules.pop(k, None)


def test_hubert_asr_run_step_bookkeeping_matches_reference():
    """HubertASR.run_step (avatars/audio_features/hubert.py:27-51): 2B chunks forwarded, silence tracking over TWO batches (features
    are computed unless this batch and the previous one were all silence), the whole l+r+2B context handed to the extractor, one
    list of B windows queued, l+r chunks kept.  In the build container the reference's own class runs the same event sequence
    (speech, then two silent steps) with its Audio2Feature replaced by a recorder; every queue / buffer / flag must match."""
    from livetalking_b200.plugin.hubert_asr import HubertASR

    class FakeFeatures:                                   # stands in for livetalking_b200.hubert.HubertFeatures
        def __init__(self, B):
            self.B, self.calls = B, []

        def run(self, pcm):
            self.calls.append(pcm.copy())
            return np.ones((self.B, 16, 1024), np.float32)

    Bsz = 3
    opt = stubs.Opt(batch_size=Bsz)

    def drive(asr):
        _feed(asr, 20 + 2 * Bsz, np.random.default_rng(4))
        asr.warm_up()
        shapes = []
        for _step in range(3):                            # speech; silence (previous was speech -> still computed); silence (skipped)
            asr.run_step()
            shapes.append([np.asarray(f).shape for f in asr.feat_queue.get()])
        return shapes

    fake = FakeFeatures(Bsz)
    ours = HubertASR(opt, None, fake, audio_feat_length=[4, 4])
    shapes = drive(ours)
    assert shapes == [[(16, 1024)] * Bsz, [(16, 1024)] * Bsz, [(10, 1024)] * Bsz]
    assert len(fake.calls) == 2 and fake.calls[0].size == (20 + 2 * Bsz) * 320 and ours.last_is_silence
    with pytest.raises(RuntimeError):
        HubertASR(opt, None, None)                        # no engine object -> loud failure, never a CPU fallback

    ref_path = "/root/reference/avatars/audio_features/hubert.py"
    if not os.path.exists(ref_path):
        return
    import types
    a2f = types.ModuleType("avatars.ultralight.audio2feature")
    a2f.Audio2Feature = object
    sys.modules.setdefault("avatars.ultralight", types.ModuleType("avatars.ultralight"))
    sys.modules["avatars.ultralight.audio2feature"] = a2f
    spec = importlib.util.spec_from_file_location("ref_base_asr3", "/root/reference/avatars/audio_features/base_asr.py")
    ref_base = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_base)
    af = types.ModuleType("avatars.audio_features")
    af.__path__ = []
    sys.modules["avatars.audio_features"] = af
    sys.modules["avatars.audio_features.base_asr"] = ref_base
    spec = importlib.util.spec_from_file_location("ref_hubert_asr", ref_path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    class Recorder:                                        # the reference's audio_processor
        def __init__(self):
            self.calls = []

        def get_hubert_from_16k_speech(self, pcm):
            self.calls.append(np.asarray(pcm).copy())
            return np.ones(((len(pcm) - 80) // 320, 1024), np.float32)

    rec = Recorder()
    theirs = ref.HubertASR(opt, None, rec, audio_feat_length=[4, 4])
    tshapes = drive(theirs)
    assert tshapes == shapes
    assert len(rec.calls) == len(fake.calls) and all(np.array_equal(a, b) for a, b in zip(rec.calls, fake.calls))
    assert theirs.last_is_silence == ours.last_is_silence
    assert len(theirs.frames) == len(ours.frames) and all(np.array_equal(x, y) for x, y in zip(theirs.frames, ours.frames))
    assert theirs.output_queue.qsize() == ours.output_queue.qsize()
    for _ in range(theirs.output_queue.qsize()):
        a, b = theirs.output_queue.get(), ours.output_queue.get()
        assert a.type == b.type and np.array_equal(a.data, b.data)
    for k in ("avatars.audio_features", "avatars.audio_features.base_asr", "avatars.ultralight.audio2feature", "avatars.ultralight"):
        sys.modules.pop(k, None)
