"""Whisper-tiny audio features on the B200 engine — replaces ``Audio2Feature.audio2feat``
(avatars/musetalk/whisper/audio2feature.py:106-117: HF ``WhisperFeatureExtractor`` + ``WhisperModel.encoder(...,
output_hidden_states=True)``) and the slicing of ``WhisperASR.run_step`` (avatars/audio_features/whisper.py:58-76).

Weights come from the HF ``WhisperModel`` state_dict (``encoder.*`` keys).  The encoder (2 conv1d + 4 pre-LN transformer
layers over 1500 steps) is assembled from the same engine ops as the UNet and captured into one CUDA graph together with
the log-mel kernels and the per-frame (50, 384) slicing."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from ._capi import check, lib
from .musetalk import Builder, _Norm, _np
from .ops import ConvWeight, Ctx, DevTensor

N_FRAMES, N_MELS, N_BINS, N_SAMPLES = 3000, 80, 201, 480000


def slaney_mel_filterbank() -> np.ndarray:
    """transformers.audio_utils.mel_filter_bank(201, 80, 0, 8000, 16000, norm='slaney', mel_scale='slaney').T -> (80, 201) f32."""
    def hz_to_mel(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-12) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)

    def mel_to_hz(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp(np.log(6.4) / 27.0 * (m - 15.0)), 200.0 * m / 3.0)

    fft_freqs = np.linspace(0, 8000, N_BINS)
    filt = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(8000.0), N_MELS + 2))
    fd = np.diff(filt)
    slopes = filt[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / fd[:-1]
    up = slopes[:, 2:] / fd[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (filt[2:N_MELS + 2] - filt[:N_MELS]))[None, :]
    return np.ascontiguousarray(fb.T.astype(np.float32))


class _WAttn:
    self_attn = True

    def __init__(self, ctx: Ctx, sd, p: str, d_model: int, heads: int):
        self.heads, self.d, self.dp = heads, d_model // heads, d_model // heads
        assert self.d % 16 == 0
        w = np.concatenate([_np(sd[f"{p}.{n}.weight"]) for n in ("q_proj", "k_proj", "v_proj")], 0)
        b = np.concatenate([_np(sd[f"{p}.q_proj.bias"]), np.zeros(d_model, np.float32), _np(sd[f"{p}.v_proj.bias"])])   # k_proj has no bias
        self.qkv = ConvWeight(ctx, w, b, tap_major=False)
        self.out = ConvWeight(ctx, _np(sd[p + ".out_proj.weight"]), _np(sd[p + ".out_proj.bias"]), tap_major=False)


class WhisperEncoder:
    def __init__(self, ctx: Ctx, sd: Dict, d_model: int = 384, heads: int = 6, layers: int = 4):
        sd = {k[len("encoder."):] if k.startswith("encoder.") else k: v for k, v in sd.items() if not k.startswith("decoder.")}
        self.ctx, self.D, self.heads, self.L = ctx, d_model, heads, layers
        # conv1d(k=3) as 1x3 convs over a (1, 1, T, C) NHWC tensor
        self.conv1 = ConvWeight(ctx, _np(sd["conv1.weight"])[:, :, None, :], _np(sd["conv1.bias"]), tap_major=False)
        self.conv2 = ConvWeight(ctx, _np(sd["conv2.weight"])[:, :, None, :], _np(sd["conv2.bias"]), tap_major=False)
        self.pos = ctx.upload(_np(sd["embed_positions.weight"]).astype(np.float16))          # (1500, D)
        self.layers = []
        for i in range(layers):
            p = f"layers.{i}"
            self.layers.append({
                "ln1": _Norm(ctx, sd, p + ".self_attn_layer_norm"), "attn": _WAttn(ctx, sd, p + ".self_attn", d_model, heads),
                "ln2": _Norm(ctx, sd, p + ".final_layer_norm"),
                "fc1": ConvWeight(ctx, _np(sd[p + ".fc1.weight"]), _np(sd[p + ".fc1.bias"]), tap_major=False),
                "fc2": ConvWeight(ctx, _np(sd[p + ".fc2.weight"]), _np(sd[p + ".fc2.bias"]), tap_major=False)})
        self.ln_post = _Norm(ctx, sd, "layer_norm")
        self.fb = ctx.upload(slaney_mel_filterbank())
        ctx.sync()

    def emit(self, b: Builder, feats16: DevTensor):
        """feats16: fp16 (3000, 80) log-mel features -> the 5 hidden states HF returns, each (1500, D) fp16.
        Ops are enqueued on the builder's ctx (the session's own stream); the weights live in self.ctx (read-only)."""
        ctx, D = b.ctx, self.D
        T = N_FRAMES
        x = DevTensor(feats16.ptr, (1, 1, T, N_MELS))
        h = b.new(1, 1, T, D)
        ctx.conv(x, self.conv1, h, N=1, IH=1, IW=T, OH=1, OW=T, pad=(0, 1))
        ctx.eltwise(h, None, T * D, 8, 1, h)                                             # GELU
        T2 = T // 2
        h2 = b.new(1, 1, T2, D)
        ctx.conv(h, self.conv2, h2, N=1, IH=1, IW=T, OH=1, OW=T2, stride=(1, 2), pad=(0, 1))
        ctx.eltwise(h2, None, T2 * D, 8, 1, h2)                                          # GELU
        x = b.new(T2, D)
        ctx.eltwise(h2, self.pos, T2 * D, T2 * D, 0, x)                                  # + embed_positions
        hidden = [x]
        for i, L in enumerate(self.layers):
            x = b.attention(L["attn"], b.layernorm(x, L["ln1"]), 1, T2, res=x)
            f = b.linear(b.layernorm(x, L["ln2"]), L["fc1"])
            ctx.eltwise(f, None, f.rows * f.C, 8, 1, f)
            x = b.linear(f, L["fc2"], res=x)
            hidden.append(x)
        hidden[-1] = b.layernorm(x, self.ln_post)                                         # HF applies the final LN to the last state
        return hidden


class WhisperFeatures:
    """audio2feat + WhisperASR slicing for one session: PCM buffer -> (B, 50, D) features, one CUDA graph."""

    def __init__(self, enc: WhisperEncoder, batch: int, stride_left: int = 10, stride_right: int = 10, out: Optional[DevTensor] = None,
                 out_rows: int = 50, keep_hidden: bool = False, ctx: Optional[Ctx] = None):
        """ctx: this extractor's own stream + scratch (created here unless given): WhisperASR.run_step runs on the render
        thread concurrently with inference_batch on the inference thread (avatars/base_avatar.py:483-489 vs :366)."""
        self.enc, self.B = enc, int(batch)
        self._own_ctx = ctx is None
        ctx = self.ctx = Ctx() if ctx is None else ctx
        self.n = (stride_left + stride_right + 2 * self.B) * 320
        if self.n > N_SAMPLES:
            raise ValueError("audio window longer than 30 s")
        self.pcm = ctx.alloc((self.n,), np.float32, zero=True)
        self.logspec = ctx.alloc((N_MELS * N_FRAMES,), np.float32, zero=True)
        self.gmax = ctx.alloc((4,), np.int32, zero=True)
        self.feats16 = ctx.alloc((N_FRAMES, N_MELS), np.float16, zero=True)
        self.feats32 = ctx.alloc((N_MELS, N_FRAMES), np.float32, zero=True) if keep_hidden else None
        self.out_rows = out_rows
        self.out = out if out is not None else ctx.alloc((self.B, out_rows, enc.D), np.float16, zero=True)
        self.start = stride_left / 2.0
        self.builder = Builder(ctx)

        def emit():
            check(lib().ltb_op_whisper_logmel(ctx._h, C.c_void_p(self.pcm.ptr), self.n, C.c_void_p(enc.fb.ptr), C.c_void_p(self.logspec.ptr),
                                              C.c_void_p(self.gmax.ptr), C.c_void_p(self.feats16.ptr),
                                              C.c_void_p(self.feats32.ptr) if self.feats32 is not None else None))
            self.hidden = enc.emit(self.builder, self.feats16)
            ptrs = (C.c_void_p * 5)(*[h.ptr for h in self.hidden])
            check(lib().ltb_op_whisper_slice(ctx._h, ptrs, N_FRAMES // 2, enc.D, self.B, float(self.start), 2.0, C.c_void_p(self.out.ptr),
                                             self.out_rows))

        emit()
        ctx.sync()
        from .musetalk import _Replay
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
        """-> (B, 50, D) float16, the list WhisperASR.run_step queues (stacked)."""
        with self.ctx.lock:
            self.run_async(pcm)
            full = self.ctx.download(self.out)
        return full[:, :50]

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
