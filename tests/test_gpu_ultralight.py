"""GPU: the UltraLight path through the C ABI — U-Net + LightReal glue + paste-back against the CPU oracle
(oracle/ultralight_ref.py, pinned to the reference modules by tests/test_ultralight_oracle.py), the HuBERT feature extractor against
the reference's own third-party implementation (transformers.HubertModel + Wav2Vec2FeatureExtractor), and the plugin classes driven
like BaseAvatar's threads."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import stubs  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ul_sd():
    from oracle import ultralight_ref as U
    return U.synth_state_dict(0)


def _avatar_assets(n, seed, H=260, W=340):
    from oracle import ultralight_ref as U
    _img, _aud, faces = U.synth_inputs(n, seed=seed)
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    # (x1,y1,x2,y2): generic stretch, identity 168x168, exact 2x decimation 84x84, tall sliver, wide
    boxes = [(30, 20, 230, 200), (10, 40, 178, 208), (200, 100, 284, 184), (5, 3, 36, 250), (0, 0, 340, 120)]
    coords = [boxes[i % len(boxes)] for i in range(n)]
    return frames, faces, coords


def test_depthwise_and_bilinear_ops_match_torch():
    """ltb_op_dwconv3x3 / ltb_op_upsample_bilinear2x on channel slices against plain PyTorch fp32 (fp16 in/out, fp32 accumulate:
    |err| <= 2e-3 + 2e-3 |ref|)."""
    import torch.nn.functional as F
    from livetalking_b200 import engine
    from livetalking_b200.ops import Ctx, DevTensor
    engine.set_device(0)
    ctx = Ctx()
    g = torch.Generator().manual_seed(3)
    for (N, H, W, C, stride) in ((2, 21, 19, 24, 1), (3, 20, 20, 64, 2), (1, 7, 9, 16, 2)):
        pitch, off = C + 16, 8
        buf = (torch.randn(N, H, W, pitch, generator=g) * 0.8).half()
        x = buf[..., off:off + C].float().permute(0, 3, 1, 2)
        w = torch.randn(C, 1, 3, 3, generator=g) * 0.4
        b = torch.randn(C, generator=g) * 0.2
        ref = F.relu(F.conv2d(x, w.half().float(), b, stride, 1, 1, C)).permute(0, 2, 3, 1).numpy()
        OH, OW = ref.shape[1], ref.shape[2]
        dbuf = ctx.upload(buf.numpy())
        dx = DevTensor(dbuf.ptr, (N, H, W, C), pitch=pitch, c_off=off)
        dw = ctx.upload(np.ascontiguousarray(w.numpy().reshape(C, 9).T).astype(np.float16))
        db = ctx.upload(b.numpy().astype(np.float32))
        opitch = C + 8
        dout = ctx.alloc((N, OH, OW, opitch), np.float16, zero=True)
        ctx.dwconv3x3(dx, N, H, W, dw, db, stride, True, DevTensor(dout.ptr, (N, OH, OW, C), pitch=opitch, c_off=8))
        got = ctx.download(dout).astype(np.float32)
        assert not got[..., :8].any()
        err = np.abs(got[..., 8:] - ref)
        assert (err <= 2e-3 + 2e-3 * np.abs(ref)).all(), (N, H, W, C, stride, err.max())
        # bilinear x2 align_corners=True into a slice
        up_ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True).permute(0, 2, 3, 1).numpy()
        dup = ctx.alloc((N, 2 * H, 2 * W, 2 * C), np.float16, zero=True)
        ctx.upsample_bilinear2x(dx, N, H, W, DevTensor(dup.ptr, (N, 2 * H, 2 * W, C), pitch=2 * C, c_off=C))
        gu = ctx.download(dup).astype(np.float32)
        assert not gu[..., :C].any()
        err = np.abs(gu[..., C:] - up_ref)
        assert (err <= 2e-3 + 2e-3 * np.abs(up_ref)).all(), ("bilinear", N, H, W, C, err.max())
    ctx.close()


def test_ultralight_unet_glue_and_paste_match_oracle(ul_sd):
    """UltraLightSession: prep (crop / rectangle mask / /255) + U-Net + head vs the oracle chain on the same crops and audio windows
    (fp16 weights with the BatchNorm scales folded in, fp16 activations, fp32 accumulate — three roundings per InvertedResidual and no
    non-linearity after the projection conv: PSNR >= 40 dB on the u8 image, stage taps within 4e-2 of max and 2.5e-2 of mean, measured
    1.8e-2 / 1.8e-2 after all 28 blocks); paste-back bit-exact against
    the oracle's OpenCV-pinned restatement applied to the engine's own prediction, for every bbox class."""
    from livetalking_b200 import engine
    from livetalking_b200.ops import Ctx
    from livetalking_b200.ultralight import UltraLightAvatar, UltraLightModel, UltraLightSession
    from oracle import ultralight_ref as U
    engine.set_device(0)
    n, B, index = 5, 4, 3                                             # frames 3,4,4,3 (mirror)
    frames, faces, coords = _avatar_assets(n, seed=12)
    _i, audio, _f = U.synth_inputs(B, seed=40)
    feats = audio.numpy().reshape(B, 16, 1024)
    taps_o = {}
    idxs = [U.mirror_index(n, index + i) for i in range(B)]
    img = torch.stack([U.lightreal_image(faces[i]) for i in idxs])
    want = (U.unet_forward(ul_sd, img, audio, taps=taps_o).numpy().transpose(0, 2, 3, 1) * 255.0)
    ctx = Ctx()
    model = UltraLightModel(ctx, ul_sd)
    av = UltraLightAvatar(ctx, model, frames, faces, coords)
    s = UltraLightSession(av, B, keep_taps=True)
    pred = s.infer(index, feats)
    assert pred.shape == (B, 160, 160, 3) and pred.dtype == np.float32 and np.isfinite(pred).all()
    img_got = s.ctx.download(s.img16).astype(np.float32)
    np.testing.assert_allclose(img_got[..., :6], img.permute(0, 2, 3, 1).numpy(), atol=5e-4)            # fp16 rounding of x/255
    assert not img_got[..., 6:].any()
    bad = []
    for k in ("x5", "audio", "fuse", "u1", "u2", "u3", "u4"):
        g = s.ctx.download_slice(s.taps[k]).astype(np.float32)
        w = taps_o[k].permute(0, 2, 3, 1).numpy()
        rel = np.abs(g - w).max() / max(1e-6, np.abs(w).max())
        mrel = np.abs(g - w).mean() / max(1e-6, np.abs(w).mean())
        if rel > 4e-2 or mrel > 2.5e-2:
            bad.append(f"{k}:{rel:.4f}:{mrel:.5f}")
    assert not bad, "taps out of tolerance (name:max-rel:mean-rel): " + "; ".join(bad)
    assert U.psnr_u8(pred.astype(np.uint8), want.astype(np.uint8)) >= 40.0, U.psnr_u8(pred.astype(np.uint8), want.astype(np.uint8))
    assert np.abs(pred - want).mean() <= 1.0 and np.abs(pred - want).max() <= 16.0          # x255 units; steepest sigmoid pixels
    # paste-back: every frame of the batch, then each bbox class from a host prediction
    allf = s.paste_batch(index)
    for i in range(B):
        assert np.array_equal(allf[i], U.lightreal_paste(pred[i], frames[idxs[i]], faces[idxs[i]], coords[idxs[i]])), i
    fused = s.infer_paste(index, feats)
    assert np.abs(fused.astype(int) - allf.astype(int)).max() <= 1
    host_pred = np.random.default_rng(1).uniform(0, 255.99, (160, 160, 3)).astype(np.float32)
    for idx in range(n):
        got = s.paste_pred(host_pred, idx)
        assert got.flags.writeable and np.array_equal(got, U.lightreal_paste(host_pred, frames[idx], faces[idx], coords[idx])), idx
    s.close()
    ctx.close()


def _hubert(layers=2, inter=1024):
    from transformers import HubertConfig, HubertModel
    torch.manual_seed(0)
    cfg = HubertConfig(hidden_size=1024, num_hidden_layers=layers, num_attention_heads=16, intermediate_size=inter, feat_extract_norm="layer",
                       conv_bias=True, do_stable_layer_norm=True)             # the layout of hubert-large-ls960-ft, fewer layers
    m = HubertModel(cfg).eval()
    with torch.no_grad():                                                     # give the random-init network some dynamic range
        for name, p in m.named_parameters():
            if name.endswith("bias"):
                p.add_(torch.randn_like(p) * 0.05)
            elif p.ndim >= 2 and "pos_conv" not in name and "feature_extractor" not in name:
                p.mul_(3.0)
    return m


def test_hubert_features_match_transformers():
    """HubertFeatures (processor normalisation + conv stack + positional conv + stable-LN encoder + window gather) against
    transformers' Wav2Vec2FeatureExtractor + HubertModel driven as Audio2Feature.get_hubert_from_16k_speech does
    (audio2feature.py:14-56) and BaseASR._feature2chunks as HubertASR.run_step calls it.  Tolerance: fp16 activations vs the fp32
    reference, |err| <= 4e-2 max|ref|, mean |err| <= 1e-2 mean|ref|."""
    from transformers import Wav2Vec2FeatureExtractor
    from livetalking_b200 import engine
    from livetalking_b200.hubert import HubertEncoder, HubertFeatures
    from livetalking_b200.ops import Ctx
    from oracle import ultralight_ref as U
    engine.set_device(0)
    model = _hubert()
    B = 8
    n = (10 + 10 + 2 * B) * 320
    rng = np.random.default_rng(7)
    t = np.arange(n) / 16000.0
    pcm = (0.3 * np.sin(2 * np.pi * 220 * t) + 0.1 * np.sin(2 * np.pi * 1900 * t) + 0.05 * rng.standard_normal(n) + 0.02).astype(np.float32)
    fe = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True, return_attention_mask=True)
    with torch.no_grad():
        hid = model(fe(pcm, return_tensors="pt", sampling_rate=16000).input_values).last_hidden_state[0].numpy()
    ref = U.trim_features(hid, n)
    rows = U.window_rows(ref.shape[0], B, 5.0)
    want = ref[rows]                                                           # (B, 16, 1024)
    ctx = Ctx()
    enc = HubertEncoder(ctx, model.state_dict())
    hf = HubertFeatures(enc, B)
    got = hf.run(pcm)
    assert got.shape == (B, 16, 1024) and got.dtype == np.float32
    h = hf.hidden_states().astype(np.float32)
    assert h.shape == hid.shape
    rel = np.abs(h - hid).max() / np.abs(hid).max()
    mrel = np.abs(h - hid).mean() / np.abs(hid).mean()
    assert rel < 4e-2 and mrel < 1e-2, (rel, mrel)
    err = np.abs(got - want)
    assert err.max() <= 4e-2 * np.abs(want).max() and err.mean() <= 1e-2 * np.abs(want).mean()
    again = hf.run(pcm)                                                        # graph replay
    assert np.array_equal(again, got)
    hf.close()
    ctx.close()


def test_lightreal_session_loop(ul_sd):
    """LightReal + HubertASR driven like BaseAvatar's threads (registry.create, put_audio_frame, run_step, inference_batch,
    paste_back_frame) in the fused mode and in the reference's data flow (ltb_return_pred), against the oracle chain on the
    engine's own HuBERT windows (the extractor has its own parity test)."""
    stubs.install()
    from livetalking_b200.plugin import ultralight_avatar as UL
    from oracle import ultralight_ref as U
    import registry
    model = UL.make_model(_hubert(layers=1, inter=512).state_dict())
    n, B = 3, 2
    frames, faces, coords = _avatar_assets(n, seed=5, H=240, W=320)
    coords = [(30, 20, 230, 200), (10, 40, 178, 208), (200, 100, 284, 184)]
    payload = UL.make_avatar(ul_sd, list(frames), list(faces), coords)
    UL.warm_up(B, payload, 160)
    rng = np.random.default_rng(8)
    t = np.arange((20 + 2 * B) * 320) / 16000.0
    audio = (0.3 * np.sin(2 * np.pi * 300 * t) + 0.05 * rng.standard_normal(t.size)).astype(np.float32)
    outs = {}
    for mode in ("fused", "pred"):
        av = registry.create("avatar", "ultralight", opt=stubs.Opt(batch_size=B, ltb_return_pred=(mode == "pred")), model=model, avatar=payload)
        for c in range(2 * B):
            av.asr.put_audio_frame(audio[c * 320:(c + 1) * 320], {})
        av.asr.run_step()
        feats = av.asr.feat_queue.get(timeout=2)
        assert len(feats) == B and feats[0].shape == (16, 1024) and feats[0].dtype == np.float32
        index = 1
        pred = av.inference_batch(index, feats)
        idxs = [U.mirror_index(n, index + i) for i in range(B)]
        out = [av.paste_back_frame(pred[i], idxs[i]) for i in range(B)]
        for o in out:
            assert o.shape == (240, 320, 3) and o.dtype == np.uint8 and o.flags.writeable
        if mode == "pred":
            assert isinstance(pred, np.ndarray) and pred.shape == (B, 160, 160, 3) and pred.dtype == np.float32
            want = U.lightreal_inference_batch(ul_sd, list(faces), index, feats)
            assert U.psnr_u8(pred.astype(np.uint8), want.astype(np.uint8)) >= 40.0
            for i in range(B):
                assert np.array_equal(out[i], U.lightreal_paste(pred[i], frames[idxs[i]], faces[idxs[i]], coords[idxs[i]]))
        else:
            with pytest.raises(ValueError):
                av.paste_back_frame(pred[0], (idxs[0] + 1) % n)
        outs[mode] = out
        av.close()
    for a, b in zip(outs["fused"], outs["pred"]):
        assert np.abs(a.astype(int) - b.astype(int)).max() <= 1
