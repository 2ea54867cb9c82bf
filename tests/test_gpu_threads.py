"""GPU: one engine session driven by three threads, as the reference's render() drives it.

Round 1 let `ltb_w2l_mel_step` (render thread) and `ltb_w2l_infer` (inference thread) share the device buffer `s->mel`
with no lock: a mel step landing between infer's H2D and its graph launch made the forward consume the NEXT batch's
windows.  These tests fail on that code:
  * hammer: mel_step in a tight loop on one thread while another thread runs infer + paste_batch; every result must be
    bit-identical to the single-threaded result for the same (index, mel windows);
  * session loop: the full plugin (MelASR + LipReal) under three threads for >= 200 frames of speech bursts + silence;
    every emitted frame must equal the frame a SECOND, single-threaded engine session produces for the same audio window
    and avatar index (the engine is deterministic: bit-exact)."""
import threading
import time

import numpy as np
import pytest

import ref_runtime as RR
import stubs

pytestmark = pytest.mark.gpu

B, N_AV, H, W = 4, 6, 120, 160


def _assets(seed=0):
    rng = np.random.default_rng(seed)
    faces = [rng.integers(0, 256, (256, 256, 3), dtype=np.uint8) for _ in range(N_AV)]
    frames = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(N_AV)]
    coords = [(10 + i, 90 + i, 20 + 2 * i, 120 + 2 * i) for i in range(N_AV)]
    return faces, frames, coords


def test_mel_step_hammer_does_not_disturb_inference(w2l_state_dict):
    from livetalking_b200 import engine
    engine.set_device(0)
    faces, frames, coords = _assets()
    model = engine.W2LModel.from_state_dict(w2l_state_dict)
    av = engine.W2LAvatar(faces, frames, coords)
    s = engine.W2LSession(model, av, B)
    rng = np.random.default_rng(1)
    n = (10 + 10 + 2 * B) * 320
    mels = [np.clip(rng.standard_normal((B, 80, 16)), -4, 4).astype(np.float32) for _ in range(6)]
    want = []
    for k, m in enumerate(mels):                                     # single-threaded ground truth
        pred = s.infer(3 * k, m)
        want.append((pred, s.paste_batch(3 * k)))
    stop = threading.Event()
    errors = []

    def hammer():
        r = np.random.default_rng(2)
        try:
            while not stop.is_set():
                s.mel_step((0.3 * r.standard_normal(n)).astype(np.float32))
        except Exception as e:                                       # noqa: BLE001
            errors.append(e)

    th = threading.Thread(target=hammer)
    th.start()
    try:
        for it in range(200):
            k = it % len(mels)
            pred = s.infer(3 * k, mels[k])
            fr = s.paste_batch(3 * k)
            assert np.array_equal(pred, want[k][0]), f"iteration {it}: forward consumed the wrong mel windows"
            assert np.array_equal(fr, want[k][1])
    finally:
        stop.set()
        th.join()
    assert not errors, errors
    s.close()
    av.close()
    model.close()


def test_lipreal_three_thread_session_loop(w2l_state_dict):
    stubs.install()
    from livetalking_b200 import engine
    from livetalking_b200.plugin import wav2lip_avatar as P
    import sys
    mirror = sys.modules["utils.image"].mirror_index
    engine.set_device(0)
    faces, frames, coords = _assets(3)
    model = engine.W2LModel.from_state_dict(w2l_state_dict)
    payload = P.make_avatar(frames, faces, coords)
    avatar = P.LipReal(stubs.Opt(batch_size=B), model, payload)
    sink = RR.RecordingSink()
    AFD = sys.modules["avatars.base_avatar"].AudioFrameData
    pulled = [AFD(data=np.zeros(320, np.float32), type=1, userdata={}) for _ in range(20)]     # warm_up() on an empty queue
    RR.spy_audio_frames(avatar.asr, pulled)
    quit_event = threading.Event()
    th = threading.Thread(target=stubs.run_three_threads, args=(avatar, sink, quit_event))
    th.start()
    feeder = threading.Thread(target=RR.feed_bursts, args=(_Feeder(avatar), [90, 70, 110, 50]))
    feeder.start()
    t0 = time.time()
    while len(sink.frames) < 220 and time.time() - t0 < 120:
        time.sleep(0.02)
    quit_event.set()
    th.join(timeout=60)
    feeder.join(timeout=30)
    assert not th.is_alive()
    n = len(sink.frames)
    assert n >= 200, f"only {n} frames emitted"
    # single-threaded replay on a second session of the same model / avatar
    ref = engine.W2LSession(model, payload.engine_avatar, B)
    index, k, j, n_speech = 0, 0, 0, 0
    while j < n:
        buf = pulled[2 * B * k: 2 * B * k + 20 + 2 * B]
        if len(buf) < 20 + 2 * B:
            break
        out = pulled[2 * B * k + 10: 2 * B * k + 10 + 2 * B]
        silent = all(c.type != 0 for c in out)
        if not silent:
            mel = ref.mel_step(np.concatenate([np.asarray(c.data, np.float32) for c in buf]))
            ref.infer(index, mel, want_pred=False)
            pasted = ref.paste_batch(index)
        for i in range(B):
            if j >= n:
                break
            idx = mirror(N_AV, index)
            if silent or all(c.type != 0 for c in out[2 * i:2 * i + 2]):
                want = frames[idx]
            else:
                want = pasted[i]
                n_speech += 1
            assert np.array_equal(sink.frames[j], want), f"frame {j} (step {k}, slot {i}) differs from the single-threaded engine"
            index += 1
            j += 1
        k += 1
    assert j >= 200 and 40 <= n_speech <= j - 20, (j, n_speech)
    ref.close()
    avatar.engine_session.close()


class _Feeder:
    """feed_bursts calls put_audio_frame(chunk, datainfo) on the avatar; the stub BaseAvatar has no such method."""

    def __init__(self, avatar):
        self.avatar = avatar

    def put_audio_frame(self, chunk, datainfo):
        self.avatar.asr.put_audio_frame(chunk, datainfo)
