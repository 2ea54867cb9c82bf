"""HuBERT audio features on the B200 engine — replaces ``Audio2Feature.get_hubert_from_16k_speech``
(avatars/ultralight/audio2feature.py:14-56: ``Wav2Vec2Processor`` + ``HubertModel(...).last_hidden_state``) and the window gather of
``HubertASR.run_step`` (avatars/audio_features/hubert.py:27-51).

Weights come from the HF ``HubertModel`` state_dict of the checkpoint the reference loads (hubert-large-ls960-ft: 7 conv layers with
per-layer LayerNorm and bias, 1024-d stable-LayerNorm transformer, 16-group positional conv with weight norm).  Conv layers 1-6, the
projections, attention and MLPs run on the tcgen05 conv / fused-attention kernels; conv layer 0 (+ the processor's utterance
normalisation), the grouped positional conv and the window gather are csrc/hubert.cu.  One CUDA graph per extractor."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .musetalk import Builder, _Norm, _Replay, _np
from .ops import ConvWeight, Ctx, DevTensor

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)
POS_K, POS_GROUPS = 128, 16
WIN = (4, 4)            # HubertASR(audio_feat_length=[4,4]) (ultralight_avatar.py:137)
ROWS = 16               # (4 + 4) * 2 feature rows per frame -> reshape(16, 32, 32)


def conv_frames(n: int) -> int:
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        n = (n - k) // s + 1
    return n


class _HAttn:
    self_attn = True

    def __init__(self, ctx: Ctx, sd, p: str, d_model: int, heads: int):
        self.heads, self.d, self.dp = heads, d_model // heads, d_model // heads
        if self.d % 16:
            raise ValueError("HuBERT head dim must be a multiple of 16")
        w = np.concatenate([_np(sd[f"{p}.{n}.weight"]) for n in ("q_proj", "k_proj", "v_proj")], 0)
        b = np.concatenate([_np(sd[f"{p}.{n}.bias"]) for n in ("q_proj", "k_proj", "v_proj")])
        self.qkv = ConvWeight(ctx, w, b, tap_major=False)
        self.out = ConvWeight(ctx, _np(sd[p + ".out_proj.weight"]), _np(sd[p + ".out_proj.bias"]), tap_major=False)


class HubertEncoder:
    """Device-resident ``HubertModel`` (large layout: feat_extract_norm='layer', conv_bias, do_stable_layer_norm)."""

    def __init__(self, ctx: Ctx, sd: Dict, heads: int = 16, eps: float = 1e-5):
        sd = {k[len("hubert."):] if k.startswith("hubert.") else k: v for k, v in sd.items()}
        if "feature_extractor.conv_layers.1.layer_norm.weight" not in sd or "feature_extractor.conv_layers.0.conv.bias" not in sd:
            raise ValueError("HubertEncoder supports the hubert-large layout (feat_extract_norm='layer', conv_bias=True) the reference loads")
        self.ctx, self.heads, self.eps = ctx, heads, eps
        fe = "feature_extractor.conv_layers"
        self.C = int(_np(sd[f"{fe}.0.conv.weight"]).shape[0])
        self.conv0_w = ctx.upload(_np(sd[f"{fe}.0.conv.weight"]).reshape(self.C, CONV_KERNEL[0]).astype(np.float32))
        self.conv0_b = ctx.upload(_np(sd[f"{fe}.0.conv.bias"]).astype(np.float32))
        # conv1d(k) as 1 x k convs over a (1, 1, T, C) NHWC tensor
        self.convs = [ConvWeight(ctx, _np(sd[f"{fe}.{i}.conv.weight"])[:, :, None, :], _np(sd[f"{fe}.{i}.conv.bias"]), tap_major=False)
                      for i in range(1, 7)]
        self.conv_ln = [_Norm(ctx, sd, f"{fe}.{i}.layer_norm") for i in range(7)]
        self.proj_ln = _Norm(ctx, sd, "feature_projection.layer_norm")
        self.proj = ConvWeight(ctx, _np(sd["feature_projection.projection.weight"]), _np(sd["feature_projection.projection.bias"]), tap_major=False)
        self.D = self.proj.cout
        pc = "encoder.pos_conv_embed.conv"
        if f"{pc}.parametrizations.weight.original0" in sd:
            g, v = _np(sd[f"{pc}.parametrizations.weight.original0"]), _np(sd[f"{pc}.parametrizations.weight.original1"])
        elif f"{pc}.weight_g" in sd:
            g, v = _np(sd[f"{pc}.weight_g"]), _np(sd[f"{pc}.weight_v"])
        else:
            g, v = None, _np(sd[f"{pc}.weight"])
        if g is not None:                                              # weight_norm(dim=2): w[:, :, k] = g[k] * v[:, :, k] / ||v[:, :, k]||
            v = v * (g / np.sqrt((v.astype(np.float64) ** 2).sum((0, 1), keepdims=True))).astype(np.float32)
        if v.shape != (self.D, self.D // POS_GROUPS, POS_K) or self.D // POS_GROUPS != 64:
            raise ValueError(f"positional conv must be ({self.D},{self.D // POS_GROUPS},{POS_K}) with 64 channels per group, got {v.shape}")
        self.pos_w = ctx.upload(np.ascontiguousarray(v.transpose(0, 2, 1)).astype(np.float16))      # [D][K][D/G]
        self.pos_b = ctx.upload(_np(sd[f"{pc}.bias"]).astype(np.float32))
        if "encoder.layers.0.attention.q_proj.weight" not in sd:
            raise ValueError("no encoder layers in the state_dict")
        self.layers = []
        i = 0
        while f"encoder.layers.{i}.attention.q_proj.weight" in sd:
            p = f"encoder.layers.{i}"
            self.layers.append({
                "ln1": _Norm(ctx, sd, p + ".layer_norm"), "attn": _HAttn(ctx, sd, p + ".attention", self.D, heads),
                "ln2": _Norm(ctx, sd, p + ".final_layer_norm"),
                "fc1": ConvWeight(ctx, _np(sd[p + ".feed_forward.intermediate_dense.weight"]), _np(sd[p + ".feed_forward.intermediate_dense.bias"]),
                                  tap_major=False),
                "fc2": ConvWeight(ctx, _np(sd[p + ".feed_forward.output_dense.weight"]), _np(sd[p + ".feed_forward.output_dense.bias"]),
                                  tap_major=False)})
            i += 1
        self.ln_post = _Norm(ctx, sd, "encoder.layer_norm")
        ctx.sync()

    def emit(self, b: Builder, pcm: DevTensor, n: int, stats: DevTensor) -> DevTensor:
        """pcm: float32 [n] raw 16 kHz samples -> last_hidden_state (conv_frames(n), D) fp16 (HubertModel.forward on the
        processor-normalised input; stable-LayerNorm encoder: hidden += pos_conv(hidden); pre-LN layers; final LayerNorm)."""
        ctx, C, D = b.ctx, self.C, self.D
        T = (n - CONV_KERNEL[0]) // CONV_STRIDE[0] + 1
        h = b.new(1, 1, T, C)
        ctx.hubert_conv0(pcm, n, self.conv0_w, self.conv0_b, C, stats, h)
        for i in range(7):
            if i > 0:
                k, s = CONV_KERNEL[i], CONV_STRIDE[i]
                T2 = (T - k) // s + 1
                h2 = b.new(1, 1, T2, C)
                ctx.conv(h, self.convs[i - 1], h2, N=1, IH=1, IW=T, OH=1, OW=T2, stride=(1, s), pad=(0, 0))
                h, T = h2, T2
            y = b.new(1, 1, T, C)
            ctx.layernorm(h, T, C, self.eps, self.conv_ln[i].gamma, self.conv_ln[i].beta, y)      # HubertLayerNormConvLayer
            ctx.eltwise(y, None, T * C, 8, 1, y)                                                  # GELU
            h = y
        x = DevTensor(h.ptr, (T, C))
        x = b.linear(b.layernorm(x, self.proj_ln, self.eps), self.proj)                           # HubertFeatureProjection
        xp = b.new(T, D)
        ctx.hubert_pos_conv(x, T, D, POS_GROUPS, POS_K, self.pos_w, self.pos_b, xp)               # + positional conv embedding
        x = xp
        for L in self.layers:                                                                     # HubertEncoderLayerStableLayerNorm
            x = b.attention(L["attn"], b.layernorm(x, L["ln1"], self.eps), 1, T, res=x)
            f = b.linear(b.layernorm(x, L["ln2"], self.eps), L["fc1"])
            ctx.eltwise(f, None, f.rows * f.C, 8, 1, f)
            x = b.linear(f, L["fc2"], res=x)
        return b.layernorm(x, self.ln_post, self.eps)


class HubertFeatures:
    """get_hubert_from_16k_speech + HubertASR's window gather for one session: PCM buffer -> (B, 16, D) features, one CUDA graph.
    The window is (stride_left + stride_right + 2 * batch) 20 ms chunks (HubertASR keeps exactly that many, hubert.py:30-48): always
    below the reference's 320000-sample clip length, so the single-clip branch of audio2feature.py:38-47 applies."""

    def __init__(self, enc: HubertEncoder, batch: int, stride_left: int = 10, stride_right: int = 10, out_nhwc: Optional[DevTensor] = None,
                 ctx: Optional[Ctx] = None):
        self.enc, self.B = enc, int(batch)
        self._own_ctx = ctx is None
        ctx = self.ctx = Ctx() if ctx is None else ctx
        self.n = (stride_left + stride_right + 2 * self.B) * 320
        if not 400 <= self.n < 320000:
            raise ValueError("audio window must be 400 .. 319999 samples")
        self.Tc, self.T = conv_frames(self.n), (self.n - 80) // 320
        if abs(self.Tc - self.T) > 1:
            raise ValueError("conv frame count and expected_T differ by more than one (audio2feature.py:52)")
        self.pcm = ctx.alloc((self.n,), np.float32, zero=True)
        self.stats = ctx.alloc((4,), np.float32, zero=True)
        self.out = ctx.alloc((self.B, ROWS, enc.D), np.float32, zero=True)
        self.out_nhwc = out_nhwc
        self.start = stride_left / 2.0
        self.builder = Builder(ctx)

        def emit():
            self.hidden = enc.emit(self.builder, self.pcm, self.n, self.stats)
            ctx.hubert_slice(self.hidden, self.Tc, self.T, enc.D, self.B, ROWS, self.start, 2.0, WIN[0], self.out, self.out_nhwc)

        emit()
        ctx.sync()
        temps, self.builder.temps = self.builder.temps, []
        self.builder.new = _Replay(temps)
        with ctx.capture() as cap:
            emit()
        self.graph = cap.graph

    def run_async(self, pcm: Optional[np.ndarray] = None):
        if pcm is not None:
            pcm = np.ascontiguousarray(pcm, np.float32).reshape(-1)
            if pcm.size != self.n:
                raise ValueError(f"expected {self.n} samples, got {pcm.size}")
            self.ctx.h2d(self.pcm, pcm, sync=False)
        self.graph.launch()

    def run(self, pcm: np.ndarray) -> np.ndarray:
        """-> (B, 16, D) float32: the list HubertASR.run_step queues (stacked)."""
        with self.ctx.lock:
            self.run_async(pcm)
            return self.ctx.download(self.out)

    def hidden_states(self) -> np.ndarray:
        with self.ctx.lock:
            return self.ctx.download(self.hidden)

    def close(self):
        if getattr(self, "graph", None) is not None:
            self.graph.close()
            self.graph = None
        if self._own_ctx and self.ctx is not None:
            self.ctx.close()
        self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gflop_per_window(n_samples: int, layers: int = 24, d_model: int = 1024, ffn: int = 4096, conv_dim: int = 512) -> float:
    """Algorithmic GFLOP (2 x MAC) of one HubertModel forward over n_samples of 16 kHz audio."""
    fl, t, cin = 0.0, n_samples, 1
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        t = (t - k) // s + 1
        fl += 2.0 * t * conv_dim * cin * k
        cin = conv_dim
    fl += 2.0 * t * conv_dim * d_model + 2.0 * t * d_model * (d_model // POS_GROUPS) * POS_K
    fl += layers * (2.0 * t * (4 * d_model * d_model + 2 * d_model * ffn) + 4.0 * t * t * d_model)
    return fl / 1e9
