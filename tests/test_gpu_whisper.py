"""GPU: Whisper audio features (log-mel + tiny encoder + per-frame slicing) against the reference's own third-party
implementation — transformers.WhisperFeatureExtractor / WhisperModel (installed in this image), driven exactly as
Audio2Feature.audio2feat and WhisperASR.run_step do (audio2feature.py:106-117, whisper.py:35-76)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(pcm, model, B, start=5.0):
    from transformers import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor()
    feats = fe(pcm, return_tensors="pt", sampling_rate=16000).input_features              # (1, 80, 3000)
    with torch.no_grad():
        hs = model.encoder(feats, output_hidden_states=True).hidden_states
    stacked = torch.stack(hs, dim=2).squeeze(0).numpy()                                     # (1500, 5, 384)
    chunks = []
    for i in range(B):                                                                      # base_asr.py:91-133 with win [0,5], mult 2
        center = int((i + start) * 2)
        idx = [min(max(j, 0), 1499) for j in range(center, center + 10)]
        chunks.append(stacked[idx].reshape(-1, 384))
    return feats[0].numpy(), [h[0].numpy() for h in hs], np.stack(chunks)


def test_whisper_features_match_transformers():
    from transformers import WhisperConfig, WhisperModel
    from livetalking_b200 import engine
    from livetalking_b200.ops import Ctx
    from livetalking_b200.whisper import WhisperEncoder, WhisperFeatures
    engine.set_device(0)
    torch.manual_seed(0)
    cfg = WhisperConfig(d_model=384, encoder_layers=4, encoder_attention_heads=6, encoder_ffn_dim=1536, decoder_layers=1,
                        decoder_attention_heads=6, decoder_ffn_dim=64, num_mel_bins=80, max_source_positions=1500)
    model = WhisperModel(cfg).eval()
    with torch.no_grad():                                   # give the random-init encoder some dynamic range
        for n, p in model.encoder.named_parameters():
            if p.ndim >= 2 and "embed_positions" not in n:
                p.mul_(4.0)
            elif n.endswith("bias"):
                p.add_(torch.randn_like(p) * 0.05)
    B = 8
    n = (10 + 10 + 2 * B) * 320
    rng = np.random.default_rng(7)
    t = np.arange(n) / 16000.0
    pcm = (0.3 * np.sin(2 * np.pi * 220 * t) + 0.1 * np.sin(2 * np.pi * 1900 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    feats_ref, hidden_ref, chunks_ref = _reference(pcm, model, B)
    ctx = Ctx()
    enc = WhisperEncoder(ctx, model.state_dict())
    wf = WhisperFeatures(enc, B, keep_hidden=True)
    got = wf.run(pcm).astype(np.float32)
    feats = ctx.download(wf.feats32)
    np.testing.assert_allclose(feats, feats_ref, atol=2e-3)
    for i, (h, hr) in enumerate(zip(wf.hidden, hidden_ref)):
        g = ctx.download(h).astype(np.float32)
        rel = np.abs(g - hr).max() / max(1e-6, np.abs(hr).max())
        mrel = np.abs(g - hr).mean() / max(1e-6, np.abs(hr).mean())
        assert rel < 4e-2 and mrel < 1e-2, (i, rel, mrel)
    assert got.shape == (B, 50, 384)
    err = np.abs(got - chunks_ref)
    assert err.max() <= 4e-2 * np.abs(chunks_ref).max() and err.mean() <= 1e-2 * np.abs(chunks_ref).mean()
    # silence: the zero-padded region sits on (max - 8 + 4) / 4
    wf.run(np.zeros(n, np.float32))
    f0 = ctx.download(wf.feats32)
    from transformers import WhisperFeatureExtractor
    np.testing.assert_allclose(f0, WhisperFeatureExtractor()(np.zeros(n, np.float32), return_tensors="np", sampling_rate=16000).input_features[0],
                               atol=1e-4)
    ctx.close()
