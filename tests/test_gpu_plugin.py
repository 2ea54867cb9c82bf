"""GPU: the plugin classes (LipReal / MelASR) driven the way BaseAvatar.inference / process_frames drive them,
checked against the CPU oracle restatement of the reference pipeline."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(__file__))
import stubs  # noqa: E402


def _pipeline_oracle(sd, faces, frames, coords, pcm, B, index):
    from oracle import mel_ref, paste_ref
    from oracle import wav2lip_ref as R
    mel = mel_ref.mel_step(pcm, B)
    img = paste_ref.w2l_build_batch(faces, index, B)
    out = R.wav2lip_forward(sd, torch.from_numpy(mel.astype(np.float32)).reshape(B, 1, 80, 16), torch.from_numpy(img))
    pred = out.numpy().transpose(0, 2, 3, 1) * 255.0
    res = []
    for i in range(B):
        idx = paste_ref.mirror_index(len(faces), index + i)
        res.append(paste_ref.w2l_paste_back(pred[i], frames[idx], coords[idx]))
    return mel, pred, res


@pytest.mark.parametrize("return_pred", [False, True])
def test_lipreal_session_loop(w2l_state_dict, return_pred):
    stubs.install()
    from livetalking_b200 import engine
    from livetalking_b200.plugin import wav2lip_avatar as W
    from oracle import wav2lip_ref as R
    import registry
    engine.set_device(0)
    B = 4
    rng = np.random.default_rng(3)
    _, img = R.synth_inputs(3, seed=21)
    faces = list((img[:, 3:6].permute(0, 2, 3, 1).numpy() * 255.0).round().astype(np.uint8))
    frames = list(rng.integers(0, 256, (3, 240, 320, 3), dtype=np.uint8))
    coords = [(10, 170, 40, 210), (20, 148, 60, 188), (0, 240, 0, 320)]
    model = engine.W2LModel.from_state_dict(w2l_state_dict)
    avatar = W.make_avatar(frames, faces, coords)
    opt = stubs.Opt(batch_size=B, ltb_return_pred=return_pred)
    av = registry.create("avatar", "wav2lip", opt=opt, model=model, avatar=avatar)      # app.py:99
    assert av.get_avatar_length() == 3 and av.asr.feat_queue.maxsize == 2
    # TTS side: 20 ms chunks of speech (tone + noise, benchmark_asr.py recipe)
    n_chunks = 2 * B
    t = np.arange((20 + n_chunks) * 320) / 16000.0
    audio = (0.3 * np.sin(2 * np.pi * 300 * t) + 0.05 * rng.standard_normal(t.size)).astype(np.float32)
    # the session warmed up on silence (20 zero chunks); now speech arrives
    for c in range(n_chunks):
        av.asr.put_audio_frame(audio[c * 320:(c + 1) * 320], {"c": c})
    av.asr.run_step()                                                                   # render thread
    feats = av.asr.feat_queue.get(timeout=1)                                            # inference thread
    chunks = [av.asr.output_queue.get() for _ in range(10 + 2 * B)][10:]
    assert all(f.type == 0 for f in chunks)
    index = 5
    pred = av.inference_batch(index, feats)
    outs = [av.paste_back_frame(p, W.mirror_index(3, index + i)) for i, p in enumerate(pred)]   # process_frames thread
    pcm = np.concatenate([np.zeros(20 * 320, np.float32), audio[:n_chunks * 320]])
    mel_o, pred_o, res_o = _pipeline_oracle(w2l_state_dict, faces, frames, coords, pcm, B, index)
    np.testing.assert_allclose(np.stack(feats), mel_o, atol=1e-5)
    for i in range(B):
        assert outs[i].dtype == np.uint8 and outs[i].shape == (240, 320, 3) and outs[i].flags.writeable and outs[i].flags.c_contiguous
        assert R.psnr_u8(outs[i], res_o[i]) >= 40.0, i
        y1, y2, x1, x2 = coords[W.mirror_index(3, index + i)]
        mask = np.ones((240, 320), bool)
        mask[y1:y2, x1:x2] = False
        assert np.array_equal(outs[i][mask], frames[W.mirror_index(3, index + i)][mask])   # outside the bbox: untouched copy
    if return_pred:
        assert isinstance(pred, np.ndarray) and pred.shape == (B, 256, 256, 3) and pred.dtype == np.float32
        assert R.psnr_u8(pred.astype(np.uint8), pred_o.astype(np.uint8)) >= 40.0
    else:
        with pytest.raises(ValueError):
            av.paste_back_frame(pred[0], (pred[0].idx + 1) % 3)
    outs[0][:] = 0                                                                       # caller may scribble (cv2.putText)
    av.engine_session.close()


def test_musereal_session_loop():
    """MuseReal + WhisperASR driven like BaseAvatar's threads, small networks, vs the CPU oracle chain."""
    stubs.install()
    from transformers import WhisperConfig, WhisperModel
    from livetalking_b200.plugin import musetalk_avatar as MT
    from oracle import musetalk_ref as M
    from oracle import paste_ref as P
    from oracle.wav2lip_ref import psnr_u8
    import registry
    torch.manual_seed(1)
    wm = WhisperModel(WhisperConfig(d_model=384, encoder_layers=4, encoder_attention_heads=6, encoder_ffn_dim=1536, decoder_layers=1,
                                    decoder_attention_heads=6, decoder_ffn_dim=64)).eval()
    us, vs = M.synth_unet_state_dict(M.UNET_SMALL), M.synth_vae_state_dict(M.VAE_SMALL)
    model = MT.make_model(us, vs, wm.state_dict(), M.UNET_SMALL, M.VAE_SMALL)
    assert len(tuple(model)) == 5                                    # the reference unpacks 5 objects from load_model()
    B, n = 2, 3
    rng = np.random.default_rng(8)
    lat, _ = M.synth_latents_and_audio(n, seed=11)
    frames = list(rng.integers(0, 256, (n, 200, 260, 3), dtype=np.uint8))
    coords = [(60, 30, 190, 170), (50, 20, 200, 180), (70, 40, 180, 160)]
    crops = [(30, 10, 230, 195), (20, 5, 240, 198), (40, 20, 220, 190)]
    masks = [np.repeat((np.linspace(0, 255, (c[3] - c[1]))[:, None] * np.ones((1, c[2] - c[0]))).astype(np.uint8)[..., None], 3, 2) for c in crops]
    avatar = MT.make_avatar(frames, masks, coords, crops, [lat[i:i + 1] for i in range(n)], model)
    av = registry.create("avatar", "musetalk", opt=stubs.Opt(batch_size=B), model=model, avatar=avatar)
    t = np.arange((20 + 2 * B) * 320) / 16000.0
    audio = (0.3 * np.sin(2 * np.pi * 300 * t) + 0.05 * rng.standard_normal(t.size)).astype(np.float32)
    for c in range(2 * B):
        av.asr.put_audio_frame(audio[c * 320:(c + 1) * 320], {})
    av.asr.run_step()
    feats = av.asr.feat_queue.get(timeout=1)
    assert len(feats) == B and feats[0].shape == (50, 384)
    index = 1
    pred = av.inference_batch(index, feats)
    assert pred.shape == (B, 256, 256, 3) and pred.dtype == np.uint8
    # oracle chain on the engine's own whisper features (the whisper path has its own parity test)
    aud = torch.from_numpy(np.stack(feats).astype(np.float32))
    idxs = [P.mirror_index(n, index + i) for i in range(B)]
    want = M.decode_latents_u8(vs, M.VAE_SMALL, M.unet_forward(us, M.UNET_SMALL, lat[idxs], M.positional_encoding(aud)))
    assert psnr_u8(pred, want) >= 40.0, psnr_u8(pred, want)
    for i in range(B):
        out = av.paste_back_frame(pred[i], idxs[i])
        assert np.array_equal(out, P.mt_paste_back(pred[i], frames[idxs[i]], coords[idxs[i]], masks[idxs[i]], crops[idxs[i]]))
        assert out.flags.writeable and out.flags.c_contiguous


def test_load_avatar_prefers_packed_assets(tmp_path, monkeypatch):
    """plugin.load_avatar: the packed avatar.ltbav yields the same payload and the same resident
    assets as the reference's directory of PNGs + coords.pkl (wav2lip_avatar.py:72-88)."""
    import pickle
    import cv2
    stubs.install()
    from livetalking_b200 import avatar_pack, engine
    from livetalking_b200.plugin import wav2lip_avatar as W
    engine.set_device(0)
    rng = np.random.default_rng(8)
    root = tmp_path / "data" / "avatars" / "demo"
    (root / "full_imgs").mkdir(parents=True)
    (root / "face_imgs").mkdir()
    coords = []
    for i in range(3):
        cv2.imwrite(str(root / "full_imgs" / f"{i:08d}.png"), rng.integers(0, 256, (96, 128, 3), dtype=np.uint8))
        cv2.imwrite(str(root / "face_imgs" / f"{i:08d}.png"), rng.integers(0, 256, (256, 256, 3), dtype=np.uint8))
        coords.append((4 + i, 84 + i, 10, 100))
    pickle.dump(coords, open(root / "coords.pkl", "wb"))
    monkeypatch.chdir(tmp_path)
    plain = W.load_avatar("demo")                                   # directory form
    avatar_pack.pack_wav2lip(str(root))
    packed = W.load_avatar("demo")                                  # packed form is picked up
    for a, b in zip(plain[0] + plain[1], packed[0] + packed[1]):
        assert np.array_equal(a, b)
    assert [tuple(c) for c in plain[2]] == [tuple(c) for c in packed[2]]
    ea, eb = plain.engine_avatar, packed.engine_avatar
    assert (ea.n, ea.H, ea.W) == (eb.n, eb.H, eb.W) == (3, 96, 128)
    assert np.array_equal(ea.faces, eb.faces) and np.array_equal(ea.frames, eb.frames) and np.array_equal(ea.coords, eb.coords)
    ea.close()
    eb.close()


def test_lipreal_cross_session_mode(w2l_state_dict):
    """opt.ltb_cross_session: two sessions with different avatars and small batch sizes submit their frames
    to the shared scheduler from two threads; each gets its own frames back, equal (PSNR >= 40 dB after the u8 conversion) to
    the CPU oracle's frames for its own audio and avatar, and the scheduler packed requests of both sessions into common launches."""
    import threading
    stubs.install()
    from livetalking_b200 import engine
    from livetalking_b200.plugin import wav2lip_avatar as W
    from oracle import wav2lip_ref as R
    import registry
    engine.set_device(0)
    B, H, Wd = 2, 240, 320
    rng = np.random.default_rng(8)
    model = engine.W2LModel.from_state_dict(w2l_state_dict)
    sessions = []
    for s in range(2):
        _, img = R.synth_inputs(3, seed=30 + s)
        faces = list((img[:, 3:6].permute(0, 2, 3, 1).numpy() * 255.0).round().astype(np.uint8))
        frames = list(rng.integers(0, 256, (3, H, Wd, 3), dtype=np.uint8))
        coords = [(10 + s, 170 + s, 40, 210), (20, 148, 60 + s, 188 + s), (0, 240, 0, 320)]
        av = registry.create("avatar", "wav2lip", opt=stubs.Opt(batch_size=B, ltb_cross_session=True, sessionid=s), model=model,
                             avatar=W.make_avatar(frames, faces, coords))
        t = np.arange((20 + 2 * B) * 320) / 16000.0
        audio = (0.3 * np.sin(2 * np.pi * (250 + 90 * s) * t) + 0.05 * rng.standard_normal(t.size)).astype(np.float32)
        sessions.append((av, faces, frames, coords, audio))
    batcher = sessions[0][0]._batcher
    assert batcher is not None and sessions[1][0]._batcher is batcher
    results = [None, None]

    def run(s):
        av, faces, frames, coords, audio = sessions[s]
        for c in range(2 * B):
            av.asr.put_audio_frame(audio[c * 320:(c + 1) * 320], {})
        av.asr.run_step()
        feats = av.asr.feat_queue.get(timeout=5)
        index = 1 + s
        outs = []
        for _rep in range(6):                                        # several steps so that requests of both sessions meet in a batch
            pred = av.inference_batch(index, feats)
            outs = [av.paste_back_frame(p, W.mirror_index(3, index + i)) for i, p in enumerate(pred)]
        results[s] = (index, outs)

    ths = [threading.Thread(target=run, args=(s,)) for s in range(2)]
    for t_ in ths:
        t_.start()
    for t_ in ths:
        t_.join(timeout=120)
    for s in range(2):
        av, faces, frames, coords, audio = sessions[s]
        index, outs = results[s]
        pcm = np.concatenate([np.zeros(20 * 320, np.float32), audio[:2 * B * 320]])
        _, _, res_o = _pipeline_oracle(w2l_state_dict, faces, frames, coords, pcm, B, index)
        for i in range(B):
            assert outs[i].shape == (H, Wd, 3) and outs[i].flags.writeable
            assert R.psnr_u8(outs[i], res_o[i]) >= 40.0, (s, i)
    assert batcher.slots == 2 * 6 * B and batcher.batches <= 2 * 6         # (packing itself is asserted deterministically in tests/test_batcher.py)
    batcher.close()
    for av, *_ in sessions:
        av.close()


def test_musereal_cross_session_mode():
    """opt.ltb_cross_session for MuseTalk: three sessions with their own avatars call inference_batch from three
    threads; the shared scheduler packs their group requests into common UNet + VAE launches (MuseTalkBatchSession) and every session
    gets ITS predictions back — equal (PSNR >= 40 dB) to the CPU oracle chain on its own latents / features — and pastes them with
    its own assets, bit-exact against the blend oracle."""
    import threading
    stubs.install()
    from transformers import WhisperConfig, WhisperModel
    from livetalking_b200.plugin import musetalk_avatar as MT
    from oracle import musetalk_ref as M
    from oracle import paste_ref as P
    from oracle.wav2lip_ref import psnr_u8
    import registry
    torch.manual_seed(1)
    wm = WhisperModel(WhisperConfig(d_model=384, encoder_layers=4, encoder_attention_heads=6, encoder_ffn_dim=1536, decoder_layers=1,
                                    decoder_attention_heads=6, decoder_ffn_dim=64)).eval()
    us, vs = M.synth_unet_state_dict(M.UNET_SMALL), M.synth_vae_state_dict(M.VAE_SMALL)
    model = MT.make_model(us, vs, wm.state_dict(), M.UNET_SMALL, M.VAE_SMALL)
    B, n, S = 2, 3, 3
    rng = np.random.default_rng(8)
    coords = [(60, 30, 190, 170), (50, 20, 200, 180), (70, 40, 180, 160)]
    crops = [(30, 10, 230, 195), (20, 5, 240, 198), (40, 20, 220, 190)]
    masks = [np.repeat((np.linspace(0, 255, (c[3] - c[1]))[:, None] * np.ones((1, c[2] - c[0]))).astype(np.uint8)[..., None], 3, 2) for c in crops]
    sessions = []
    for s in range(S):
        lat, aud = M.synth_latents_and_audio(n, seed=20 + s)
        frames = list(rng.integers(0, 256, (n, 200, 260 + 4 * s, 3), dtype=np.uint8))
        avatar = MT.make_avatar(frames, masks, coords, crops, [lat[i:i + 1] for i in range(n)], model)
        av = registry.create("avatar", "musetalk", opt=stubs.Opt(batch_size=B, ltb_cross_session=True, sessionid=s), model=model, avatar=avatar)
        sessions.append((av, lat, aud[:B], frames))
    batcher = sessions[0][0]._batcher
    assert batcher is not None and all(x[0]._batcher is batcher for x in sessions)
    results = [None] * S

    def run(s):
        av, lat, aud, frames = sessions[s]
        feats = [aud[i].numpy().astype(np.float32) for i in range(B)]
        for _rep in range(4):
            pred = av.inference_batch(s, feats)
        results[s] = pred

    ths = [threading.Thread(target=run, args=(s,)) for s in range(S)]
    for t_ in ths:
        t_.start()
    for t_ in ths:
        t_.join(timeout=180)
    for s in range(S):
        av, lat, aud, frames = sessions[s]
        pred = results[s]
        assert pred is not None and pred.shape == (B, 256, 256, 3) and pred.dtype == np.uint8
        idxs = [P.mirror_index(n, s + i) for i in range(B)]
        want = M.decode_latents_u8(vs, M.VAE_SMALL, M.unet_forward(us, M.UNET_SMALL, lat[idxs], M.positional_encoding(aud)))
        assert psnr_u8(pred, want) >= 40.0, (s, psnr_u8(pred, want))
        for i in range(B):
            out = av.paste_back_frame(pred[i], idxs[i])
            assert np.array_equal(out, P.mt_paste_back(pred[i], frames[idxs[i]], coords[idxs[i]], masks[idxs[i]], crops[idxs[i]]))
    assert batcher.slots == S * 4 and batcher.batches <= S * 4
    batcher.close()
    batcher.mux.close()
    for av, *_ in sessions:
        av.close()
