"""Synthetic (random-init) wav2lip256 weights and avatar assets in the reference's formats, for benchmarking and
bring-up when no checkpoint / avatar directory is available (there is no network in the build environment).

The state_dict uses the reference checkpoint's key scheme (avatars/wav2lip_avatar.py:59-70); BN layers are
identity (gamma=1, beta=0, mean=0, var=1) and residual branches are damped so activations stay in fp16 range.
Timing does not depend on the weight values (no data-dependent control flow anywhere in the path)."""
from __future__ import annotations

import numpy as np

from .w2l_pack import layer_table

_RES = {  # prefixes of residual Conv2d blocks (wav2lip_v2.py: residual=True)
}


def random_state_dict(seed: int = 0):
    rng = np.random.default_rng(seed)
    sd = {}
    for prefix, kind, ci, co, k in layer_table():
        shape = (co, ci, k, k) if kind == "c" else (ci, co, k, k)
        fan_in = ci * k * k if kind == "c" else ci * k * k / 4.0
        residual = (kind == "c" and ci == co and k == 3 and not prefix.endswith(".0") and "output_block" not in prefix)
        std = np.sqrt(2.0 / fan_in) * (0.3 if residual else 1.0)
        sd[f"{prefix}.conv_block.0.weight"] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(std))
        sd[f"{prefix}.conv_block.0.bias"] = (rng.standard_normal(co, dtype=np.float32) * np.float32(0.02))
        sd[f"{prefix}.conv_block.1.weight"] = np.ones(co, np.float32)
        sd[f"{prefix}.conv_block.1.bias"] = np.zeros(co, np.float32)
        sd[f"{prefix}.conv_block.1.running_mean"] = np.zeros(co, np.float32)
        sd[f"{prefix}.conv_block.1.running_var"] = np.ones(co, np.float32)
    sd["output_block.1.weight"] = rng.standard_normal((3, 32, 1, 1), dtype=np.float32) * np.float32(0.3)
    sd["output_block.1.bias"] = np.zeros(3, np.float32)
    return sd


def synthetic_avatar(n: int = 64, H: int = 720, W: int = 1280, bbox=(200, 520, 480, 800), seed: int = 0):
    """seeded faces (n,256,256,3) u8, full frames (n,H,W,3) u8, coords (n,4) = (y1,y2,x1,x2)."""
    rng = np.random.default_rng(seed)
    low = rng.integers(0, 256, (n, 16, 16, 3)).astype(np.float32)
    faces = np.clip(np.kron(low, np.ones((1, 16, 16, 1), np.float32)) + rng.integers(-8, 9, (n, 256, 256, 3)), 0, 255).astype(np.uint8)
    base = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    frames = np.empty((n, H, W, 3), np.uint8)
    for i in range(n):
        frames[i] = np.roll(base, i * 3, axis=1)
    coords = np.tile(np.asarray(bbox, np.int32), (n, 1))
    return faces, frames, coords


def sine_audio(seconds: float = 60.0, freq: float = 440.0, amp: float = 0.5, sr: int = 16000) -> np.ndarray:
    """BASELINE.json synthetic audio: amp*sin(2 pi f t), 16 kHz float32."""
    t = np.arange(int(seconds * sr), dtype=np.float64) / sr
    return (amp * np.sin(2 * np.pi * freq * t)).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------------------
# random-init MuseTalk networks (diffusers / HF key schemes) for benchmarking — values from a rotating random pool
# ----------------------------------------------------------------------------------------------------------------------
class _Rng:
    """Direct generator (same interface as _Pool) for the large HuBERT weights: 315 M values in a few seconds."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def randn(self, shape, std):
        return self.rng.standard_normal(shape, dtype=np.float32) * np.float32(std)


class _Pool:
    def __init__(self, seed):
        rng = np.random.default_rng(seed)
        self.pool = rng.standard_normal(1 << 22, dtype=np.float32)
        self.off = 0

    def randn(self, shape, std):
        n = int(np.prod(shape))
        self.off = (self.off * 31 + 977) % (self.pool.size - 1)
        return (np.resize(self.pool[self.off:], n).reshape(shape) * np.float32(std)).astype(np.float32)


def _resnet(sd, p, cin, cout, g, temb=None):
    sd[p + ".norm1.weight"], sd[p + ".norm1.bias"] = np.ones(cin, np.float32), np.zeros(cin, np.float32)
    sd[p + ".conv1.weight"], sd[p + ".conv1.bias"] = g.randn((cout, cin, 3, 3), (2.0 / (cin * 9)) ** 0.5), np.zeros(cout, np.float32)
    if temb:
        sd[p + ".time_emb_proj.weight"], sd[p + ".time_emb_proj.bias"] = g.randn((cout, temb), 0.5 / temb ** 0.5), np.zeros(cout, np.float32)
    sd[p + ".norm2.weight"], sd[p + ".norm2.bias"] = np.ones(cout, np.float32), np.zeros(cout, np.float32)
    sd[p + ".conv2.weight"], sd[p + ".conv2.bias"] = g.randn((cout, cout, 3, 3), 0.7 / (cout * 9) ** 0.5), np.zeros(cout, np.float32)
    if cin != cout:
        sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"] = g.randn((cout, cin, 1, 1), 1.0 / cin ** 0.5), np.zeros(cout, np.float32)


def _transformer(sd, p, c, ctx_dim, g):
    sd[p + ".norm.weight"], sd[p + ".norm.bias"] = np.ones(c, np.float32), np.zeros(c, np.float32)
    for n in ("proj_in", "proj_out"):
        sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"] = g.randn((c, c, 1, 1), (0.5 if n == "proj_out" else 1.0) / c ** 0.5), np.zeros(c, np.float32)
    b = p + ".transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        sd[f"{b}.{n}.weight"], sd[f"{b}.{n}.bias"] = np.ones(c, np.float32), np.zeros(c, np.float32)
    for a, kd in (("attn1", c), ("attn2", ctx_dim)):
        sd[f"{b}.{a}.to_q.weight"] = g.randn((c, c), 1.0 / c ** 0.5)
        sd[f"{b}.{a}.to_k.weight"] = g.randn((c, kd), 1.0 / kd ** 0.5)
        sd[f"{b}.{a}.to_v.weight"] = g.randn((c, kd), 1.0 / kd ** 0.5)
        sd[f"{b}.{a}.to_out.0.weight"], sd[f"{b}.{a}.to_out.0.bias"] = g.randn((c, c), 0.5 / c ** 0.5), np.zeros(c, np.float32)
    sd[b + ".ff.net.0.proj.weight"], sd[b + ".ff.net.0.proj.bias"] = g.randn((8 * c, c), 1.0 / c ** 0.5), np.zeros(8 * c, np.float32)
    sd[b + ".ff.net.2.weight"], sd[b + ".ff.net.2.bias"] = g.randn((c, 4 * c), 0.7 / (4 * c) ** 0.5), np.zeros(c, np.float32)


def random_unet_state_dict(cfg, seed: int = 0):
    g = _Pool(seed)
    sd = {}
    boc = cfg.block_out_channels
    td = boc[0] * 4
    sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"] = g.randn((td, boc[0]), 1 / boc[0] ** 0.5), np.zeros(td, np.float32)
    sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"] = g.randn((td, td), 1 / td ** 0.5), np.zeros(td, np.float32)
    sd["conv_in.weight"], sd["conv_in.bias"] = g.randn((boc[0], cfg.in_channels, 3, 3), 1 / (cfg.in_channels * 9) ** 0.5), np.zeros(boc[0], np.float32)
    skip, cin = [boc[0]], boc[0]
    for i, c in enumerate(boc):
        for j in range(cfg.layers_per_block):
            _resnet(sd, f"down_blocks.{i}.resnets.{j}", cin, c, g, td)
            if cfg.down_has_attn[i]:
                _transformer(sd, f"down_blocks.{i}.attentions.{j}", c, cfg.cross_attention_dim, g)
            cin = c
            skip.append(c)
        if i < len(boc) - 1:
            sd[f"down_blocks.{i}.downsamplers.0.conv.weight"] = g.randn((c, c, 3, 3), 1 / (c * 9) ** 0.5)
            sd[f"down_blocks.{i}.downsamplers.0.conv.bias"] = np.zeros(c, np.float32)
            skip.append(c)
    _resnet(sd, "mid_block.resnets.0", boc[-1], boc[-1], g, td)
    _transformer(sd, "mid_block.attentions.0", boc[-1], cfg.cross_attention_dim, g)
    _resnet(sd, "mid_block.resnets.1", boc[-1], boc[-1], g, td)
    cin = boc[-1]
    for i, c in enumerate(reversed(boc)):
        for j in range(cfg.layers_per_block + 1):
            _resnet(sd, f"up_blocks.{i}.resnets.{j}", cin + skip.pop(), c, g, td)
            if cfg.up_has_attn[i]:
                _transformer(sd, f"up_blocks.{i}.attentions.{j}", c, cfg.cross_attention_dim, g)
            cin = c
        if i < len(boc) - 1:
            sd[f"up_blocks.{i}.upsamplers.0.conv.weight"] = g.randn((c, c, 3, 3), 1 / (c * 9) ** 0.5)
            sd[f"up_blocks.{i}.upsamplers.0.conv.bias"] = np.zeros(c, np.float32)
    sd["conv_norm_out.weight"], sd["conv_norm_out.bias"] = np.ones(boc[0], np.float32), np.zeros(boc[0], np.float32)
    sd["conv_out.weight"], sd["conv_out.bias"] = g.randn((cfg.out_channels, boc[0], 3, 3), 1 / (boc[0] * 9) ** 0.5), np.zeros(cfg.out_channels, np.float32)
    return sd


def _vae_mid(sd, p, c, g):
    _resnet(sd, p + ".resnets.0", c, c, g)
    a = p + ".attentions.0"
    sd[a + ".group_norm.weight"], sd[a + ".group_norm.bias"] = np.ones(c, np.float32), np.zeros(c, np.float32)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        sd[f"{a}.{n}.weight"], sd[f"{a}.{n}.bias"] = g.randn((c, c), (0.5 if n == "to_out.0" else 1.0) / c ** 0.5), np.zeros(c, np.float32)
    _resnet(sd, p + ".resnets.1", c, c, g)


def random_vae_state_dict(cfg, seed: int = 1):
    g = _Pool(seed)
    sd = {}
    boc, L = cfg.block_out_channels, cfg.latent_channels
    sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"] = g.randn((boc[0], 3, 3, 3), 1 / 27 ** 0.5), np.zeros(boc[0], np.float32)
    cin = boc[0]
    for i, c in enumerate(boc):
        for j in range(cfg.layers_per_block):
            _resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", cin, c, g)
            cin = c
        if i < len(boc) - 1:
            sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = g.randn((c, c, 3, 3), 1 / (c * 9) ** 0.5)
            sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = np.zeros(c, np.float32)
    _vae_mid(sd, "encoder.mid_block", boc[-1], g)
    sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"] = np.ones(boc[-1], np.float32), np.zeros(boc[-1], np.float32)
    sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"] = g.randn((2 * L, boc[-1], 3, 3), 1 / (boc[-1] * 9) ** 0.5), np.zeros(2 * L, np.float32)
    sd["quant_conv.weight"], sd["quant_conv.bias"] = g.randn((2 * L, 2 * L, 1, 1), 1 / (2 * L) ** 0.5), np.zeros(2 * L, np.float32)
    sd["post_quant_conv.weight"], sd["post_quant_conv.bias"] = g.randn((L, L, 1, 1), 1 / L ** 0.5), np.zeros(L, np.float32)
    rev = list(reversed(boc))
    sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"] = g.randn((rev[0], L, 3, 3), 1 / (L * 9) ** 0.5), np.zeros(rev[0], np.float32)
    _vae_mid(sd, "decoder.mid_block", rev[0], g)
    cin = rev[0]
    for i, c in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            _resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", cin, c, g)
            cin = c
        if i < len(rev) - 1:
            sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = g.randn((c, c, 3, 3), 1 / (c * 9) ** 0.5)
            sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = np.zeros(c, np.float32)
    sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"] = np.ones(rev[-1], np.float32), np.zeros(rev[-1], np.float32)
    sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"] = g.randn((3, rev[-1], 3, 3), 0.6 / (rev[-1] * 9) ** 0.5), np.zeros(3, np.float32)
    return sd


def random_whisper_state_dict(d_model=384, layers=4, ffn=1536, seed: int = 2):
    g = _Pool(seed)
    sd = {"encoder.conv1.weight": g.randn((d_model, 80, 3), 1 / 240 ** 0.5), "encoder.conv1.bias": np.zeros(d_model, np.float32),
          "encoder.conv2.weight": g.randn((d_model, d_model, 3), 1 / (3 * d_model) ** 0.5), "encoder.conv2.bias": np.zeros(d_model, np.float32),
          "encoder.embed_positions.weight": g.randn((1500, d_model), 0.1),
          "encoder.layer_norm.weight": np.ones(d_model, np.float32), "encoder.layer_norm.bias": np.zeros(d_model, np.float32)}
    for i in range(layers):
        p = f"encoder.layers.{i}"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[f"{p}.self_attn.{n}.weight"] = g.randn((d_model, d_model), 1 / d_model ** 0.5)
            if n != "k_proj":
                sd[f"{p}.self_attn.{n}.bias"] = np.zeros(d_model, np.float32)
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"] = np.ones(d_model, np.float32), np.zeros(d_model, np.float32)
        sd[f"{p}.fc1.weight"], sd[f"{p}.fc1.bias"] = g.randn((ffn, d_model), 1 / d_model ** 0.5), np.zeros(ffn, np.float32)
        sd[f"{p}.fc2.weight"], sd[f"{p}.fc2.bias"] = g.randn((d_model, ffn), 1 / ffn ** 0.5), np.zeros(d_model, np.float32)
    return sd


def synthetic_musetalk_avatar(n: int = 16, H: int = 720, W: int = 1280, bbox=(480, 200, 800, 520), hw: int = 32, seed: int = 0):
    """frames, bbox (x1,y1,x2,y2), crop box = get_crop_box(expand 1.5), Gaussian-ish lower-half mask, latents."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    frames = np.stack([np.roll(base, 3 * i, axis=1) for i in range(n)])
    x1, y1, x2, y2 = bbox
    xc, yc = (x1 + x2) // 2, (y1 + y2) // 2
    s = int(max(x2 - x1, y2 - y1) // 2 * 1.5)
    crop = (max(0, xc - s), max(0, yc - s), min(W, xc + s), min(H, yc + s))
    mh, mw = crop[3] - crop[1], crop[2] - crop[0]
    yy = np.linspace(0, 1, mh)[:, None] * np.ones((1, mw))
    m = (np.clip((yy - 0.45) * 5, 0, 1) * 255).astype(np.uint8)
    mask = np.stack([m, m, m], -1)
    latents = [(rng.standard_normal((1, 8, hw, hw)) * 0.8).astype(np.float16) for _ in range(n)]
    return frames, [mask] * n, [bbox] * n, [crop] * n, latents


def random_hubert_state_dict(layers: int = 24, d_model: int = 1024, ffn: int = 4096, seed: int = 3):
    """HF ``HubertModel`` key scheme of hubert-large-ls960-ft (the checkpoint avatars/ultralight/audio2feature.py:9-10 loads):
    7 conv layers (512 ch, per-layer LayerNorm, bias), projection 512 -> d_model, weight-normed 16-group positional conv (k 128),
    `layers` stable-LayerNorm encoder layers."""
    g = _Rng(seed)
    ones, zeros = (lambda n: np.ones(n, np.float32)), (lambda n: np.zeros(n, np.float32))
    sd = {}
    for i, k in enumerate((10, 3, 3, 3, 3, 2, 2)):
        cin = 1 if i == 0 else 512
        p = f"feature_extractor.conv_layers.{i}"
        sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"] = g.randn((512, cin, k), (2.0 / (cin * k)) ** 0.5), g.randn((512,), 0.02)
        sd[f"{p}.layer_norm.weight"], sd[f"{p}.layer_norm.bias"] = ones(512), zeros(512)
    sd["feature_projection.layer_norm.weight"], sd["feature_projection.layer_norm.bias"] = ones(512), zeros(512)
    sd["feature_projection.projection.weight"], sd["feature_projection.projection.bias"] = g.randn((d_model, 512), 512 ** -0.5), zeros(d_model)
    pc = "encoder.pos_conv_embed.conv"
    v = g.randn((d_model, d_model // 16, 128), (1.0 / (128 * d_model // 16)) ** 0.5)
    sd[f"{pc}.parametrizations.weight.original1"] = v
    sd[f"{pc}.parametrizations.weight.original0"] = np.sqrt((v.astype(np.float64) ** 2).sum((0, 1), keepdims=True)).astype(np.float32)
    sd[f"{pc}.bias"] = zeros(d_model)
    sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"] = ones(d_model), zeros(d_model)
    proto = {}                                   # one layer's worth of random values, reused by every layer (benchmark weights)
    for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
        proto[n] = g.randn((d_model, d_model), d_model ** -0.5)
    proto["fc1"], proto["fc2"] = g.randn((ffn, d_model), d_model ** -0.5), g.randn((d_model, ffn), ffn ** -0.5)
    for i in range(layers):
        p = f"encoder.layers.{i}"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[f"{p}.attention.{n}.weight"], sd[f"{p}.attention.{n}.bias"] = proto[n], zeros(d_model)
        for n in ("layer_norm", "final_layer_norm"):
            sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"] = ones(d_model), zeros(d_model)
        sd[f"{p}.feed_forward.intermediate_dense.weight"], sd[f"{p}.feed_forward.intermediate_dense.bias"] = proto["fc1"], zeros(ffn)
        sd[f"{p}.feed_forward.output_dense.weight"], sd[f"{p}.feed_forward.output_dense.bias"] = proto["fc2"], zeros(d_model)
    return sd


def random_ultralight_state_dict(seed: int = 4):
    """``Model(6, 'hubert').state_dict()`` key scheme (avatars/ultralight/unet.py:184-206), He-normal convs, identity-ish BatchNorms."""
    g = _Rng(seed)
    sd = {}
    ch = [32, 64, 128, 256, 512]

    def bn(p, c):
        sd[p + ".weight"], sd[p + ".bias"] = np.ones(c, np.float32), g.randn((c,), 0.05)
        sd[p + ".running_mean"], sd[p + ".running_var"] = np.zeros(c, np.float32), np.ones(c, np.float32)

    def ir(p, inp, oup):
        hid = inp * 2
        sd[p + ".conv.0.weight"] = g.randn((hid, inp, 1, 1), (2.0 / inp) ** 0.5)
        bn(p + ".conv.1", hid)
        sd[p + ".conv.3.weight"] = g.randn((hid, 1, 3, 3), (2.0 / 9) ** 0.5)
        bn(p + ".conv.4", hid)
        sd[p + ".conv.6.weight"] = g.randn((oup, hid, 1, 1), (0.5 / hid) ** 0.5)
        bn(p + ".conv.7", oup)

    def dc(p, i, o):
        ir(p + ".double_conv.0", i, o)
        ir(p + ".double_conv.1", o, o)

    a = "audio_model"
    ir(a + ".conv1", 16, ch[1]), ir(a + ".conv2", ch[1], ch[2]), ir(a + ".conv4", ch[3], ch[3]), ir(a + ".conv6", ch[4], ch[4]), ir(a + ".conv7", ch[4], ch[4])
    for name, cin, cout in ((a + ".conv3", ch[2], ch[3]), (a + ".conv5", ch[3], ch[4])):
        sd[name + ".weight"], sd[name + ".bias"] = g.randn((cout, cin, 3, 3), (2.0 / (9 * cin)) ** 0.5), g.randn((cout,), 0.05)
    bn(a + ".bn3", ch[3]), bn(a + ".bn5", ch[4])
    dc("fuse_conv.0", ch[4] * 2, ch[4]), dc("fuse_conv.1", ch[4], ch[3])
    ir("inc.inconv.0", 6, ch[0])
    for i in range(4):
        dc(f"down{i + 1}.maxpool_conv.0", ch[i], ch[i + 1])
    for i, (ci, co) in enumerate(((ch[4], ch[3] // 2), (ch[3], ch[2] // 2), (ch[2], ch[1] // 2), (ch[1], ch[0]))):
        dc(f"up{i + 1}.conv", ci, co)
    sd["outc.conv.weight"], sd["outc.conv.bias"] = g.randn((3, ch[0], 1, 1), (2.0 / ch[0]) ** 0.5), g.randn((3,), 0.1)
    return sd


def synthetic_ultralight_avatar(n: int = 16, H: int = 720, W: int = 1280, bbox=(500, 180, 780, 460), seed: int = 0):
    """frames (n,H,W,3) uint8, 168x168 face crops, bbox (x1,y1,x2,y2) per frame (ultralight_avatar.py:63-82 on-disk content)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    frames = np.stack([np.roll(base, 3 * i, axis=1) for i in range(n)])
    faces = rng.integers(0, 256, (n, 168, 168, 3), dtype=np.uint8)
    return frames, faces, [tuple(bbox)] * n
