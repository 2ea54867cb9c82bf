"""CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py): the MuseTalk networks in plain fp32 PyTorch.

What the reference pins (and what this file restates):
    avatars/musetalk_avatar.py:130-152      MuseReal.inference_batch: latents by mirror index, PE, UNet(t=0), VAE decode
    avatars/musetalk/models/unet.py:12-27   PositionalEncoding(d_model=384)
    avatars/musetalk/models/vae.py:96-108   decode_latents: 1/scaling_factor, decode, (x/2+0.5).clamp(0,1), *255 round, RGB->BGR
    avatars/musetalk/models/vae.py:51-94    preprocess_img (BGR->RGB, /255, upper-half mask, Normalize(.5,.5)) + encode (x scaling_factor)
    avatars/musetalk/models/vae.py:110-122  get_latents_for_unet: cat(masked latents, reference latents)
    avatars/musetalk/utils/utils.py:140-175 get_image_pred: the full enc -> UNet -> dec chain (latent_dist.mode())

What it does NOT pin — **parity unpinned, architecture [NOT IN REFERENCE]**: the arithmetic of
``diffusers.UNet2DConditionModel`` / ``diffusers.AutoencoderKL`` lives in the un-vendored, unpinned ``diffusers``
package (requirements.txt:41) and MuseTalk's ``musetalk.json`` is absent from the tree; diffusers is not installed in
this image either.  The layouts below restate the published SD-1.x / sd-vae-ft-mse architectures (SURVEY.md Appendix C;
a parameter counter over the same layout reproduces the published totals) with the diffusers state_dict key scheme, so
real ``unet.pth`` / ``sd-vae`` weights load by name.  Everything here is config driven (channel widths) so the parity
tests can run a narrow network in seconds and the full-width one once.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 8
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 384
    num_heads: int = 8            # diffusers legacy "attention_head_dim = 8" means 8 heads in SD-1.x configs
    norm_groups: int = 32
    norm_eps: float = 1e-5
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)   # CrossAttnDownBlock2D x3, DownBlock2D
    up_has_attn: Tuple[bool, ...] = (False, True, True, True)     # UpBlock2D, CrossAttnUpBlock2D x3


@dataclass(frozen=True)
class VAEConfig:
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    norm_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215


UNET_FULL = UNetConfig()
VAE_FULL = VAEConfig()
UNET_SMALL = UNetConfig(block_out_channels=(64, 128, 256, 256))
VAE_SMALL = VAEConfig(block_out_channels=(32, 64, 128, 128))


# ------------------------------------------------------------------------------------------------ shared blocks
def _gn(sd, p, x, groups, eps):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def _conv(sd, p, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def resnet_block(sd, p, x, groups, eps, temb: Optional[torch.Tensor]):
    """diffusers ResnetBlock2D: norm1-silu-conv1 (+time_emb_proj(silu(emb))) norm2-silu-conv2, 1x1 shortcut if cin!=cout."""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, groups, eps)), padding=1)
    if temb is not None:
        h = h + _lin(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, groups, eps)), padding=1)
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x)
    return x + h


def attention(sd, p, x, ctx, heads):
    """diffusers Attention: softmax(q k^T / sqrt(d)) v ; to_q/k/v (bias optional), to_out.0 with bias."""
    q, k, v = _lin(sd, p + ".to_q", x), _lin(sd, p + ".to_k", ctx), _lin(sd, p + ".to_v", ctx)
    B, N, C = q.shape
    d = C // heads
    q = q.view(B, N, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2) * (d ** -0.5), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    return _lin(sd, p + ".to_out.0", o)


def transformer_2d(sd, p, x, ctx, heads, groups):
    """diffusers Transformer2DModel (conv projections) with one BasicTransformerBlock (self, cross, GEGLU FF)."""
    B, C, H, W = x.shape
    res = x
    h = _conv(sd, p + ".proj_in", _gn(sd, p + ".norm", x, groups, 1e-6))
    t = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    b = p + ".transformer_blocks.0"
    n = F.layer_norm(t, (C,), sd[b + ".norm1.weight"], sd[b + ".norm1.bias"], 1e-5)
    t = t + attention(sd, b + ".attn1", n, n, heads)
    n = F.layer_norm(t, (C,), sd[b + ".norm2.weight"], sd[b + ".norm2.bias"], 1e-5)
    t = t + attention(sd, b + ".attn2", n, ctx, heads)
    n = F.layer_norm(t, (C,), sd[b + ".norm3.weight"], sd[b + ".norm3.bias"], 1e-5)
    g = _lin(sd, b + ".ff.net.0.proj", n)
    a, gate = g.chunk(2, dim=-1)
    t = t + _lin(sd, b + ".ff.net.2", a * F.gelu(gate))
    h = t.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return _conv(sd, p + ".proj_out", h) + res


# ------------------------------------------------------------------------------------------------ UNet
def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


@torch.no_grad()
def unet_forward(sd: Dict[str, torch.Tensor], cfg: UNetConfig, latents: torch.Tensor, ctx: torch.Tensor,
                 timestep: int = 0, taps: Optional[dict] = None) -> torch.Tensor:
    """latents (B,8,h,w), ctx (B,50,384) [already positional-encoded] -> (B,4,h,w).  timestep 0: musetalk_avatar.py:61."""
    boc = cfg.block_out_channels
    G, eps, heads = cfg.norm_groups, cfg.norm_eps, cfg.num_heads
    temb = timestep_embedding(torch.tensor([timestep]), boc[0])
    temb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", temb)))
    temb = temb.expand(latents.shape[0], -1)
    h = _conv(sd, "conv_in", latents, padding=1)
    skips = [h]
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block):
            h = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", h, G, eps, temb)
            if cfg.down_has_attn[i]:
                h = transformer_2d(sd, f"down_blocks.{i}.attentions.{j}", h, ctx, heads, G)
            skips.append(h)
        if i < len(boc) - 1:
            h = _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=1)
            skips.append(h)
        if taps is not None:
            taps[f"down{i}"] = h
    h = resnet_block(sd, "mid_block.resnets.0", h, G, eps, temb)
    h = transformer_2d(sd, "mid_block.attentions.0", h, ctx, heads, G)
    h = resnet_block(sd, "mid_block.resnets.1", h, G, eps, temb)
    if taps is not None:
        taps["mid"] = h
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", h, G, eps, temb)
            if cfg.up_has_attn[i]:
                h = transformer_2d(sd, f"up_blocks.{i}.attentions.{j}", h, ctx, heads, G)
        if i < len(boc) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", h, padding=1)
        if taps is not None:
            taps[f"up{i}"] = h
    h = F.silu(_gn(sd, "conv_norm_out", h, G, eps))
    return _conv(sd, "conv_out", h, padding=1)


# ------------------------------------------------------------------------------------------------ VAE
def _vae_mid(sd, p, h, G, eps):
    h = resnet_block(sd, p + ".resnets.0", h, G, eps, None)
    B, C, H, W = h.shape
    a = p + ".attentions.0"
    t = _gn(sd, a + ".group_norm", h, G, eps).permute(0, 2, 3, 1).reshape(B, H * W, C)
    t = attention(sd, a, t, t, 1)
    h = h + t.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return resnet_block(sd, p + ".resnets.1", h, G, eps, None)


@torch.no_grad()
def vae_decode(sd: Dict[str, torch.Tensor], cfg: VAEConfig, z: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """AutoencoderKL.decode(z).sample : (B,4,h,w) -> (B,3,8h,8w) in ~[-1,1] (RGB)."""
    G, eps = cfg.norm_groups, cfg.norm_eps
    rev = tuple(reversed(cfg.block_out_channels))
    h = _conv(sd, "post_quant_conv", z)
    h = _conv(sd, "decoder.conv_in", h, padding=1)
    h = _vae_mid(sd, "decoder.mid_block", h, G, eps)
    if taps is not None:
        taps["dec_mid"] = h
    for i in range(len(rev)):
        for j in range(cfg.layers_per_block + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, G, eps, None)
        if i < len(rev) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", h, padding=1)
        if taps is not None:
            taps[f"dec_up{i}"] = h
    h = F.silu(_gn(sd, "decoder.conv_norm_out", h, G, eps))
    return _conv(sd, "decoder.conv_out", h, padding=1)


@torch.no_grad()
def vae_encode_mean(sd: Dict[str, torch.Tensor], cfg: VAEConfig, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.encode(x).latent_dist.mode() : (B,3,H,W) in [-1,1] RGB -> (B,4,H/8,W/8)."""
    G, eps = cfg.norm_groups, cfg.norm_eps
    boc = cfg.block_out_channels
    h = _conv(sd, "encoder.conv_in", x, padding=1)
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, G, eps, None)
        if i < len(boc) - 1:
            h = F.pad(h, (0, 1, 0, 1))                     # diffusers Downsample2D(padding=0): asymmetric pad
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", h, stride=2)
    h = _vae_mid(sd, "encoder.mid_block", h, G, eps)
    h = F.silu(_gn(sd, "encoder.conv_norm_out", h, G, eps))
    h = _conv(sd, "encoder.conv_out", h, padding=1)
    moments = _conv(sd, "quant_conv", h)
    return moments[:, : cfg.latent_channels]


def decode_latents_u8(sd, cfg: VAEConfig, latents: torch.Tensor):
    """vae.py:96-108 -> uint8 (B,H,W,3) BGR."""
    img = vae_decode(sd, cfg, latents / cfg.scaling_factor)
    img = (img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).float().numpy()
    img = (img * 255).round().astype("uint8")
    return img[..., ::-1]


def preprocess_img(img_bgr_u8, half_mask: bool) -> torch.Tensor:
    """vae.py:51-82 for an in-memory (H,W,3) uint8 BGR image -> (1,3,H,W) float in [-1,1] RGB."""
    import numpy as np
    x = np.asarray(img_bgr_u8)[..., ::-1].astype(np.float64) / 255.0
    x = torch.FloatTensor(np.ascontiguousarray(np.transpose(x, (2, 0, 1))))
    if half_mask:
        m = torch.zeros(x.shape[1:])
        m[: x.shape[1] // 2] = 1
        x = x * (m > 0.5)
    return ((x - 0.5) / 0.5).unsqueeze(0)


def latents_for_unet(sd, cfg: VAEConfig, img_bgr_u8) -> torch.Tensor:
    """vae.py:110-122 with latent_dist.mode() (the deterministic form used by utils.py:154): (1,8,h,w)."""
    masked = cfg.scaling_factor * vae_encode_mean(sd, cfg, preprocess_img(img_bgr_u8, True))
    ref = cfg.scaling_factor * vae_encode_mean(sd, cfg, preprocess_img(img_bgr_u8, False))
    return torch.cat([masked, ref], dim=1)


def positional_encoding(x: torch.Tensor) -> torch.Tensor:
    """avatars/musetalk/models/unet.py:12-27 : x (B,T,384) + sinusoid table."""
    B, T, D = x.shape
    pe = torch.zeros(T, D)
    pos = torch.arange(0, T, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, D, 2).float() * (-math.log(10000.0) / D))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return x + pe.unsqueeze(0)


# ------------------------------------------------------------------------------------------------ synthetic weights
class FastGen:
    """Cheap deterministic pseudo-random source for the full-width networks (850 M parameters): tensors are filled from
    a fixed 4 M-element normal pool at rotating offsets instead of drawing every value (torch.randn takes minutes)."""

    def __init__(self, seed: int):
        g = torch.Generator().manual_seed(seed)
        self.pool = torch.randn(1 << 22, generator=g)
        self.upool = torch.rand(1 << 22, generator=g)
        self.off = 0

    def _take(self, pool, shape):
        n = 1
        for s in shape:
            n *= s
        self.off = (self.off * 31 + 977) % (pool.numel() - 1)
        reps = (n + self.off + pool.numel() - 1) // pool.numel() + 1
        return pool.repeat(reps)[self.off:self.off + n].reshape(shape).clone()

    def randn(self, shape):
        return self._take(self.pool, tuple(shape) if not isinstance(shape, int) else (shape,))

    def rand(self, shape):
        return self._take(self.upool, tuple(shape) if not isinstance(shape, int) else (shape,))


def _randn(shape, g):
    return g.randn(shape) if isinstance(g, FastGen) else torch.randn(shape, generator=g)


def _rand(shape, g):
    return g.rand(shape) if isinstance(g, FastGen) else torch.rand(shape, generator=g)


def _init(shape, fan_in, g, gain=1.0):
    return _randn(shape, g) * (gain / math.sqrt(fan_in))


def _add_norm(sd, p, c, g):
    sd[p + ".weight"] = _rand(c, g) * 0.4 + 0.8
    sd[p + ".bias"] = _randn(c, g) * 0.1


def _add_conv(sd, p, cin, cout, k, g, gain=1.0):
    sd[p + ".weight"] = _init((cout, cin, k, k), cin * k * k, g, gain)
    sd[p + ".bias"] = _randn(cout, g) * 0.02


def _add_lin(sd, p, cin, cout, g, bias=True, gain=1.0):
    sd[p + ".weight"] = _init((cout, cin), cin, g, gain)
    if bias:
        sd[p + ".bias"] = _randn(cout, g) * 0.02


def _add_resnet(sd, p, cin, cout, g, temb_dim=None):
    _add_norm(sd, p + ".norm1", cin, g)
    _add_conv(sd, p + ".conv1", cin, cout, 3, g, 1.4)
    if temb_dim:
        _add_lin(sd, p + ".time_emb_proj", temb_dim, cout, g, gain=0.5)
    _add_norm(sd, p + ".norm2", cout, g)
    _add_conv(sd, p + ".conv2", cout, cout, 3, g, 0.7)
    if cin != cout:
        _add_conv(sd, p + ".conv_shortcut", cin, cout, 1, g)


def _add_transformer(sd, p, c, ctx_dim, g):
    _add_norm(sd, p + ".norm", c, g)
    _add_conv(sd, p + ".proj_in", c, c, 1, g)
    b = p + ".transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        _add_norm(sd, f"{b}.{n}", c, g)
    for a, kd in (("attn1", c), ("attn2", ctx_dim)):
        _add_lin(sd, f"{b}.{a}.to_q", c, c, g, bias=False, gain=1.5)
        _add_lin(sd, f"{b}.{a}.to_k", kd, c, g, bias=False, gain=1.5)
        _add_lin(sd, f"{b}.{a}.to_v", kd, c, g, bias=False)
        _add_lin(sd, f"{b}.{a}.to_out.0", c, c, g, gain=0.5)
    _add_lin(sd, b + ".ff.net.0.proj", c, 8 * c, g)
    _add_lin(sd, b + ".ff.net.2", 4 * c, c, g, gain=0.7)
    _add_conv(sd, p + ".proj_out", c, c, 1, g, 0.5)


def synth_unet_state_dict(cfg: UNetConfig, seed: int = 0, fast: bool = False) -> Dict[str, torch.Tensor]:
    g = FastGen(seed) if fast else torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    boc = cfg.block_out_channels
    tdim = boc[0] * 4
    _add_lin(sd, "time_embedding.linear_1", boc[0], tdim, g)
    _add_lin(sd, "time_embedding.linear_2", tdim, tdim, g)
    _add_conv(sd, "conv_in", cfg.in_channels, boc[0], 3, g)
    skip_ch = [boc[0]]
    cin = boc[0]
    for i, c in enumerate(boc):
        for j in range(cfg.layers_per_block):
            _add_resnet(sd, f"down_blocks.{i}.resnets.{j}", cin, c, g, tdim)
            if cfg.down_has_attn[i]:
                _add_transformer(sd, f"down_blocks.{i}.attentions.{j}", c, cfg.cross_attention_dim, g)
            cin = c
            skip_ch.append(c)
        if i < len(boc) - 1:
            _add_conv(sd, f"down_blocks.{i}.downsamplers.0.conv", c, c, 3, g)
            skip_ch.append(c)
    _add_resnet(sd, "mid_block.resnets.0", boc[-1], boc[-1], g, tdim)
    _add_transformer(sd, "mid_block.attentions.0", boc[-1], cfg.cross_attention_dim, g)
    _add_resnet(sd, "mid_block.resnets.1", boc[-1], boc[-1], g, tdim)
    rev = list(reversed(boc))
    cin = boc[-1]
    for i, c in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            sk = skip_ch.pop()
            _add_resnet(sd, f"up_blocks.{i}.resnets.{j}", cin + sk, c, g, tdim)
            if cfg.up_has_attn[i]:
                _add_transformer(sd, f"up_blocks.{i}.attentions.{j}", c, cfg.cross_attention_dim, g)
            cin = c
        if i < len(rev) - 1:
            _add_conv(sd, f"up_blocks.{i}.upsamplers.0.conv", c, c, 3, g)
    _add_norm(sd, "conv_norm_out", boc[0], g)
    _add_conv(sd, "conv_out", boc[0], cfg.out_channels, 3, g, 1.0)
    return sd


def _add_vae_mid(sd, p, c, g):
    _add_resnet(sd, p + ".resnets.0", c, c, g)
    a = p + ".attentions.0"
    _add_norm(sd, a + ".group_norm", c, g)
    for n in ("to_q", "to_k", "to_v"):
        _add_lin(sd, f"{a}.{n}", c, c, g, gain=1.5 if n != "to_v" else 1.0)
    _add_lin(sd, a + ".to_out.0", c, c, g, gain=0.5)
    _add_resnet(sd, p + ".resnets.1", c, c, g)


def synth_vae_state_dict(cfg: VAEConfig, seed: int = 1, fast: bool = False) -> Dict[str, torch.Tensor]:
    g = FastGen(seed) if fast else torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    boc = cfg.block_out_channels
    L = cfg.latent_channels
    # encoder
    _add_conv(sd, "encoder.conv_in", 3, boc[0], 3, g)
    cin = boc[0]
    for i, c in enumerate(boc):
        for j in range(cfg.layers_per_block):
            _add_resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", cin, c, g)
            cin = c
        if i < len(boc) - 1:
            _add_conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3, g)
    _add_vae_mid(sd, "encoder.mid_block", boc[-1], g)
    _add_norm(sd, "encoder.conv_norm_out", boc[-1], g)
    _add_conv(sd, "encoder.conv_out", boc[-1], 2 * L, 3, g)
    _add_conv(sd, "quant_conv", 2 * L, 2 * L, 1, g)
    # decoder
    _add_conv(sd, "post_quant_conv", L, L, 1, g)
    rev = list(reversed(boc))
    _add_conv(sd, "decoder.conv_in", L, rev[0], 3, g)
    _add_vae_mid(sd, "decoder.mid_block", rev[0], g)
    cin = rev[0]
    for i, c in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            _add_resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", cin, c, g)
            cin = c
        if i < len(rev) - 1:
            _add_conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3, g)
    _add_norm(sd, "decoder.conv_norm_out", rev[-1], g)
    _add_conv(sd, "decoder.conv_out", rev[-1], 3, 3, g, 0.6)
    return sd


def synth_latents_and_audio(batch: int, hw: int = 32, seed: int = 2):
    """latents ~ N(0,1) (B,8,hw,hw) rounded to fp16 (the reference stores fp16 latents), whisper-like
    features (B,50,384) rounded to fp16."""
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(batch, 8, hw, hw, generator=g).half().float()
    aud = (torch.randn(batch, 50, 384, generator=g) * 0.8).half().float()
    return lat, aud


def count_params(sd) -> int:
    return int(sum(v.numel() for v in sd.values()))
