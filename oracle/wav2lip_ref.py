"""CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py): wav2lip256 forward.

A functional fp32 restatement of the reference network

    avatars/wav2lip/models/wav2lip_v2.py:8-91   (topology)
    avatars/wav2lip/models/wav2lip_v2.py:123-163 (forward: audio enc -> face enc
                                                  -> decoder with skip concat -> head)
    avatars/wav2lip/models/conv.py:5-19          (Conv2d  = conv -> BN(eval) -> [+x] -> ReLU)
    avatars/wav2lip/models/conv.py:33-44         (Conv2dTranspose = convT -> BN(eval) -> ReLU)

driven directly by a reference-format ``state_dict`` (same key names as the
checkpoint loaded at avatars/wav2lip_avatar.py:59-70).  It is pinned against the
unmodified reference nn.Module by tests/golden/make_golden.py (run in the build
container where /root/reference exists) -> tests/golden/w2l_golden.npz.

Also: ``synth_state_dict`` — seeded, *conditioned* synthetic weights (SURVEY H1:
no checkpoint exists offline; default-init collapses the output to ~0.5).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, conv.py:10,38

# (kind, cin, cout, k, stride, pad, out_pad, residual)
#   kind: "c" = Conv2d block, "t" = Conv2dTranspose block
Spec = Tuple[str, int, int, int, Tuple[int, int], int, int, bool]


def _c(cin, cout, k, s, p, res=False) -> Spec:
    s = (s, s) if isinstance(s, int) else tuple(s)
    return ("c", cin, cout, k, s, p, 0, res)


def _t(cin, cout, k, s, p, op=0) -> Spec:
    return ("t", cin, cout, k, (s, s), p, op, False)


# wav2lip_v2.py:12-39
FACE_ENCODER: List[List[Spec]] = [
    [_c(6, 16, 7, 1, 3)],
    [_c(16, 32, 3, 2, 1), _c(32, 32, 3, 1, 1, True), _c(32, 32, 3, 1, 1, True)],
    [_c(32, 64, 3, 2, 1), _c(64, 64, 3, 1, 1, True), _c(64, 64, 3, 1, 1, True), _c(64, 64, 3, 1, 1, True)],
    [_c(64, 128, 3, 2, 1), _c(128, 128, 3, 1, 1, True), _c(128, 128, 3, 1, 1, True)],
    [_c(128, 256, 3, 2, 1), _c(256, 256, 3, 1, 1, True), _c(256, 256, 3, 1, 1, True)],
    [_c(256, 512, 3, 2, 1), _c(512, 512, 3, 1, 1, True)],
    [_c(512, 512, 3, 2, 1), _c(512, 512, 3, 1, 1, True)],
    [_c(512, 512, 4, 1, 0), _c(512, 512, 1, 1, 0)],
]

# wav2lip_v2.py:41-58
AUDIO_ENCODER: List[Spec] = [
    _c(1, 32, 3, 1, 1), _c(32, 32, 3, 1, 1, True), _c(32, 32, 3, 1, 1, True),
    _c(32, 64, 3, (3, 1), 1), _c(64, 64, 3, 1, 1, True), _c(64, 64, 3, 1, 1, True),
    _c(64, 128, 3, 3, 1), _c(128, 128, 3, 1, 1, True), _c(128, 128, 3, 1, 1, True),
    _c(128, 256, 3, (3, 2), 1), _c(256, 256, 3, 1, 1, True),
    _c(256, 512, 3, 1, 0), _c(512, 512, 1, 1, 0),
]

# wav2lip_v2.py:60-87
FACE_DECODER: List[List[Spec]] = [
    [_c(512, 512, 1, 1, 0)],
    [_t(1024, 512, 4, 1, 0), _c(512, 512, 3, 1, 1, True)],
    [_t(1024, 512, 3, 2, 1, 1), _c(512, 512, 3, 1, 1, True)],
    [_t(1024, 512, 3, 2, 1, 1), _c(512, 512, 3, 1, 1, True), _c(512, 512, 3, 1, 1, True)],
    [_t(768, 384, 3, 2, 1, 1), _c(384, 384, 3, 1, 1, True), _c(384, 384, 3, 1, 1, True)],
    [_t(512, 256, 3, 2, 1, 1), _c(256, 256, 3, 1, 1, True), _c(256, 256, 3, 1, 1, True)],
    [_t(320, 128, 3, 2, 1, 1), _c(128, 128, 3, 1, 1, True), _c(128, 128, 3, 1, 1, True)],
    [_t(160, 64, 3, 2, 1, 1), _c(64, 64, 3, 1, 1, True), _c(64, 64, 3, 1, 1, True)],
]

# wav2lip_v2.py:89-91 : Conv2d(80,32,3,1,1) ; nn.Conv2d(32,3,1) ; Sigmoid
OUTPUT_BLOCK0: Spec = _c(80, 32, 3, 1, 1)


def layer_list() -> List[Tuple[str, Spec]]:
    """(state_dict prefix, spec) for the 54 Conv/ConvT+BN blocks in execution order
    (audio encoder, face encoder, decoder, output_block.0).  ``output_block.1`` is the
    bare 1x1 conv 32->3 (+sigmoid)."""
    out = []
    for i, s in enumerate(AUDIO_ENCODER):
        out.append((f"audio_encoder.{i}", s))
    for b, blk in enumerate(FACE_ENCODER):
        for j, s in enumerate(blk):
            out.append((f"face_encoder_blocks.{b}.{j}", s))
    for b, blk in enumerate(FACE_DECODER):
        for j, s in enumerate(blk):
            out.append((f"face_decoder_blocks.{b}.{j}", s))
    out.append(("output_block.0", OUTPUT_BLOCK0))
    return out


def _block(sd: Dict[str, torch.Tensor], prefix: str, spec: Spec, x: torch.Tensor,
           taps: Optional[dict] = None) -> torch.Tensor:
    kind, cin, cout, k, s, p, op, res = spec
    w = sd[f"{prefix}.conv_block.0.weight"]
    b = sd[f"{prefix}.conv_block.0.bias"]
    if kind == "c":
        y = F.conv2d(x, w, b, stride=s, padding=p)
    else:
        y = F.conv_transpose2d(x, w, b, stride=s, padding=p, output_padding=op)
    y = F.batch_norm(y, sd[f"{prefix}.conv_block.1.running_mean"], sd[f"{prefix}.conv_block.1.running_var"],
                     sd[f"{prefix}.conv_block.1.weight"], sd[f"{prefix}.conv_block.1.bias"],
                     training=False, eps=BN_EPS)
    if res:
        y = y + x
    y = F.relu(y)
    if taps is not None:
        taps[prefix] = y
    return y


@torch.no_grad()
def wav2lip_forward(sd: Dict[str, torch.Tensor], mel: torch.Tensor, img: torch.Tensor,
                    taps: Optional[dict] = None, return_logits: bool = False) -> torch.Tensor:
    """mel (B,1,80,16) f32, img (B,6,256,256) f32 in [0,1] -> (B,3,256,256) in (0,1).

    Follows Wav2Lip.forward, wav2lip_v2.py:123-163 (4-D input branch).  ``taps`` (optional
    dict) receives every block's post-ReLU output keyed by its state_dict prefix, plus
    "logits" (pre-sigmoid)."""
    x = mel
    for i, s in enumerate(AUDIO_ENCODER):                       # :132
        x = _block(sd, f"audio_encoder.{i}", s, x, taps)
    audio_embedding = x
    feats = []
    x = img
    for b, blk in enumerate(FACE_ENCODER):                      # :136-140
        for j, s in enumerate(blk):
            x = _block(sd, f"face_encoder_blocks.{b}.{j}", s, x, taps)
        feats.append(x)
    x = audio_embedding
    for b, blk in enumerate(FACE_DECODER):                      # :142-152
        for j, s in enumerate(blk):
            x = _block(sd, f"face_decoder_blocks.{b}.{j}", s, x, taps)
        x = torch.cat((x, feats[-1]), dim=1)                    # :146  [decoder_out, skip]
        feats.pop()
    x = _block(sd, "output_block.0", OUTPUT_BLOCK0, x, taps)    # :154
    logits = F.conv2d(x, sd["output_block.1.weight"], sd["output_block.1.bias"])
    if taps is not None:
        taps["logits"] = logits
    if return_logits:
        return logits
    return torch.sigmoid(logits)


# --------------------------------------------------------------------------------------
# synthetic, conditioned weights
# --------------------------------------------------------------------------------------

def _state_shapes() -> List[Tuple[str, Tuple[int, ...]]]:
    shapes = []
    for prefix, (kind, cin, cout, k, s, p, op, res) in layer_list():
        wshape = (cout, cin, k, k) if kind == "c" else (cin, cout, k, k)
        shapes.append((f"{prefix}.conv_block.0.weight", wshape))
        shapes.append((f"{prefix}.conv_block.0.bias", (cout,)))
        shapes.append((f"{prefix}.conv_block.1.weight", (cout,)))
        shapes.append((f"{prefix}.conv_block.1.bias", (cout,)))
        shapes.append((f"{prefix}.conv_block.1.running_mean", (cout,)))
        shapes.append((f"{prefix}.conv_block.1.running_var", (cout,)))
        shapes.append((f"{prefix}.conv_block.1.num_batches_tracked", ()))
    shapes.append(("output_block.1.weight", (3, 32, 1, 1)))
    shapes.append(("output_block.1.bias", (3,)))
    return shapes


def synth_inputs(batch: int, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Seeded synthetic network inputs: smooth faces (low-passed noise, lower half of the
    first 3 channels zeroed as wav2lip_avatar.py:127-130 does) and mel-like windows in [-4,4]."""
    g = torch.Generator().manual_seed(1000 + seed)
    low = torch.rand(batch, 3, 32, 32, generator=g)
    face = F.interpolate(low, size=(256, 256), mode="bilinear", align_corners=False)
    face = (face + 0.05 * torch.rand(batch, 3, 256, 256, generator=g)).clamp(0, 1)
    face = torch.round(face * 255.0) / 255.0
    masked = face.clone()
    masked[:, :, 128:] = 0
    img = torch.cat([masked, face], dim=1)
    mel = (torch.rand(batch, 1, 80, 16, generator=g) * 8.0 - 4.0)
    return mel.float(), img.float()


@torch.no_grad()
def synth_state_dict(seed: int = 0, calib_batch: int = 2) -> Dict[str, torch.Tensor]:
    """Seeded conditioned weights in the reference checkpoint's key scheme.

    He-normal conv weights, small random conv biases, BN gamma~U(0.7,1.3), beta~N(0.1,0.3),
    and BN running statistics *calibrated* to the batch statistics seen on a synthetic
    calibration batch (layer by layer, in execution order), so activations stay O(1)
    through all 55 layers and the sigmoid output spans (0,1).  Deterministic for a given
    torch build (CPU RNG)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in _state_shapes():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(1, dtype=torch.long)
        elif name.endswith("conv_block.0.weight") or name == "output_block.1.weight":
            if "conv_block.0.weight" in name:
                prefix = name[: -len(".conv_block.0.weight")]
                spec = dict(layer_list())[prefix]
                kind, cin, cout, k, s, p, op, res = spec
                fan_in = cin * k * k if kind == "c" else cin * k * k / (s[0] * s[1])
            else:
                fan_in = 32
            sd[name] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif name.endswith("conv_block.0.bias") or name == "output_block.1.bias":
            sd[name] = torch.randn(shape, generator=g) * 0.05
        elif name.endswith("conv_block.1.weight"):
            sd[name] = torch.rand(shape, generator=g) * 0.6 + 0.7
        elif name.endswith("conv_block.1.bias"):
            sd[name] = torch.randn(shape, generator=g) * 0.3 + 0.1
        elif name.endswith("running_mean"):
            sd[name] = torch.zeros(shape)
        elif name.endswith("running_var"):
            sd[name] = torch.ones(shape)
    # residual blocks: shrink gamma so relu(BN(conv(x)) + x) does not blow up
    for prefix, spec in layer_list():
        if spec[7]:
            sd[f"{prefix}.conv_block.1.weight"] *= 0.5
            sd[f"{prefix}.conv_block.1.bias"] *= 0.5

    # calibration: set running stats = batch stats of the pre-BN conv output
    mel, img = synth_inputs(calib_batch, seed=seed + 77)

    def calib_block(prefix, spec, x):
        kind, cin, cout, k, s, p, op, res = spec
        w, b = sd[f"{prefix}.conv_block.0.weight"], sd[f"{prefix}.conv_block.0.bias"]
        y = F.conv2d(x, w, b, stride=s, padding=p) if kind == "c" else \
            F.conv_transpose2d(x, w, b, stride=s, padding=p, output_padding=op)
        n = y.numel() // y.shape[1]
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False) if n > 1 else torch.ones_like(mean)
        if n < 64:  # tiny maps (1x1..): keep stats tame
            var = var + 0.25 * (mean * mean + 1.0)
        sd[f"{prefix}.conv_block.1.running_mean"] = mean.clone()
        sd[f"{prefix}.conv_block.1.running_var"] = var.clamp_min(1e-3).clone()
        return _block(sd, prefix, spec, x)

    x = mel
    for i, s in enumerate(AUDIO_ENCODER):
        x = calib_block(f"audio_encoder.{i}", s, x)
    emb = x
    feats = []
    x = img
    for b, blk in enumerate(FACE_ENCODER):
        for j, s in enumerate(blk):
            x = calib_block(f"face_encoder_blocks.{b}.{j}", s, x)
        feats.append(x)
    x = emb
    for b, blk in enumerate(FACE_DECODER):
        for j, s in enumerate(blk):
            x = calib_block(f"face_decoder_blocks.{b}.{j}", s, x)
        x = torch.cat((x, feats.pop()), dim=1)
    x = calib_block("output_block.0", OUTPUT_BLOCK0, x)
    logits = F.conv2d(x, sd["output_block.1.weight"], sd["output_block.1.bias"])
    # scale the head so logits have std ~2 and zero mean -> sigmoid spans (0,1)
    scale = 2.0 / float(logits.std().clamp_min(1e-6))
    sd["output_block.1.weight"] = sd["output_block.1.weight"] * scale
    sd["output_block.1.bias"] = (sd["output_block.1.bias"] - logits.mean(dim=(0, 2, 3))) * scale
    return sd


def psnr_u8(a, b) -> float:
    """PSNR (peak 255) between two arrays already converted with the reference's own
    float->u8 rule (truncation for wav2lip, wav2lip_avatar.py:145)."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    mse = float(np.mean((a - b) ** 2))
    if mse == 0:
        return float("inf")
    return 10.0 * math.log10(255.0 * 255.0 / mse)
