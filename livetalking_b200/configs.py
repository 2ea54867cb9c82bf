"""Network configurations of the MuseTalk path (product-side; mirrors the fields of diffusers' config JSONs that
``UNet2DConditionModel(**musetalk.json)`` / ``AutoencoderKL.from_pretrained(sd-vae)`` consume —
avatars/musetalk/models/unet.py:36-38, vae.py:24)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 8
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 384
    num_heads: int = 8
    norm_groups: int = 32
    norm_eps: float = 1e-5
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    up_has_attn: Tuple[bool, ...] = (False, True, True, True)


@dataclass(frozen=True)
class VAEConfig:
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    norm_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215


def unet_config_from_json(j: dict) -> UNetConfig:
    heads = j.get("attention_head_dim", 8)
    if isinstance(heads, (list, tuple)):
        heads = heads[0]
    return UNetConfig(
        in_channels=j.get("in_channels", 8), out_channels=j.get("out_channels", 4),
        block_out_channels=tuple(j.get("block_out_channels", (320, 640, 1280, 1280))), layers_per_block=j.get("layers_per_block", 2),
        cross_attention_dim=j.get("cross_attention_dim", 384), num_heads=int(heads), norm_groups=j.get("norm_num_groups", 32),
        norm_eps=j.get("norm_eps", 1e-5),
        down_has_attn=tuple(t.startswith("CrossAttn") for t in j.get("down_block_types", ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",))),
        up_has_attn=tuple(t.startswith("CrossAttn") for t in j.get("up_block_types", ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3)))


def vae_config_from_json(j: dict) -> VAEConfig:
    return VAEConfig(block_out_channels=tuple(j.get("block_out_channels", (128, 256, 512, 512))), layers_per_block=j.get("layers_per_block", 2),
                     latent_channels=j.get("latent_channels", 4), norm_groups=j.get("norm_num_groups", 32),
                     scaling_factor=j.get("scaling_factor", 0.18215))
