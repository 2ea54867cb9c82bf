"""TEST INFRASTRUCTURE — CPU restatement of the UltraLight path.  Not imported by the product.

What it restates (reference file:line):
* the U-Net `Model(6, 'hubert')`                       — avatars/ultralight/unet.py:7-226 (InvertedResidual, DoubleConvDW, Up,
                                                           AudioConvHubert, Model.forward), functional, straight from a state_dict
* `LightReal.inference_batch` glue                      — avatars/ultralight_avatar.py:141-169 (crop [4:164], black rectangle
                                                           (5,5,150,145), /255, 6-channel concat, (16,32,32) audio reshape, x255)
* `LightReal.paste_back_frame`                          — avatars/ultralight_avatar.py:171-184
* the HuBERT front end around `HubertModel`             — avatars/ultralight/audio2feature.py:14-56 (processor normalisation, the
                                                           expected_T trim / pad) and the window gather of
                                                           avatars/audio_features/base_asr.py:91-157 as HubertASR.run_step calls it
                                                           (avatars/audio_features/hubert.py:42-45: win [4,4], start l/2, multiplier 2)
Pinned: tests/golden/ultralight_golden.npz is produced by tests/golden/make_golden.py from the UNMODIFIED reference modules
(unet.py imported by path; LightReal built without __init__) — tests/test_ultralight_oracle.py checks this file against it.
The HuBERT network itself is `transformers.HubertModel` (the reference's own dependency, installed): tests use it directly."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

CH = [32, 64, 128, 256, 512]          # unet.py:188


# ------------------------------------------------------------------------------------------------ structure
def ir_list() -> List[Tuple[str, int, int, int, bool, int]]:
    """Every InvertedResidual of Model(6,'hubert') in state_dict order: (prefix, inp, oup, stride, residual, expand)."""
    out = []

    def dc(prefix, cin, cout, stride):     # DoubleConvDW, unet.py:39-50
        out.append((prefix + ".double_conv.0", cin, cout, stride, False, 2))
        out.append((prefix + ".double_conv.1", cout, cout, 1, True, 2))

    out.append(("audio_model.conv1", 16, CH[1], 1, False, 2))
    out.append(("audio_model.conv2", CH[1], CH[2], 1, False, 2))
    out.append(("audio_model.conv4", CH[3], CH[3], 1, True, 2))
    out.append(("audio_model.conv6", CH[4], CH[4], 1, True, 2))
    out.append(("audio_model.conv7", CH[4], CH[4], 1, True, 2))
    dc("fuse_conv.0", CH[4] * 2, CH[4], 1)
    dc("fuse_conv.1", CH[4], CH[3], 1)
    out.append(("inc.inconv.0", 6, CH[0], 1, False, 2))
    for i, (a, b) in enumerate(((CH[0], CH[1]), (CH[1], CH[2]), (CH[2], CH[3]), (CH[3], CH[4]))):
        dc(f"down{i + 1}.maxpool_conv.0", a, b, 2)
    for i, (a, b) in enumerate(((CH[4], CH[3] // 2), (CH[3], CH[2] // 2), (CH[2], CH[1] // 2), (CH[1], CH[0]))):
        dc(f"up{i + 1}.conv", a, b, 1)
    return out


def _bn(sd, p, x, calib):
    if calib:                                           # synth_state_dict: running stats := batch stats of the calibration batch
        sd[p + ".running_mean"] = x.mean((0, 2, 3)).detach()
        sd[p + ".running_var"] = x.var((0, 2, 3), unbiased=False).detach().clamp_min(1e-4)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def _ir(sd, p, x, stride, res, calib=False):
    """InvertedResidual, unet.py:7-37: 1x1 -> BN -> ReLU -> depthwise 3x3 (stride) -> BN -> ReLU -> 1x1 -> BN (+ x)."""
    h = F.relu(_bn(sd, p + ".conv.1", F.conv2d(x, sd[p + ".conv.0.weight"]), calib))
    h = F.relu(_bn(sd, p + ".conv.4", F.conv2d(h, sd[p + ".conv.3.weight"], None, stride, 1, 1, h.shape[1]), calib))
    h = _bn(sd, p + ".conv.7", F.conv2d(h, sd[p + ".conv.6.weight"]), calib)
    return x + h if res else h


def _dc(sd, p, x, stride, calib):
    return _ir(sd, p + ".double_conv.1", _ir(sd, p + ".double_conv.0", x, stride, False, calib), 1, True, calib)


def _up(sd, p, x1, x2, calib):
    """Up.forward, unet.py:81-90: bilinear x2 (align_corners=True), centre pad to the skip's size, cat([x1, x2]), DoubleConvDW."""
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return _dc(sd, p + ".conv", torch.cat([x1, x2], 1), 1, calib)


def audio_forward(sd, a, calib=False, taps=None):
    """AudioConvHubert.forward, unet.py:143-181: (B,16,32,32) -> (B,512,10,10)."""
    p = "audio_model"
    a = _ir(sd, p + ".conv1", a, 1, False, calib)
    a = _ir(sd, p + ".conv2", a, 1, False, calib)
    a = F.relu(_bn(sd, p + ".bn3", F.conv2d(a, sd[p + ".conv3.weight"], sd[p + ".conv3.bias"], 2, 1), calib))
    a = _ir(sd, p + ".conv4", a, 1, True, calib)
    a = F.relu(_bn(sd, p + ".bn5", F.conv2d(a, sd[p + ".conv5.weight"], sd[p + ".conv5.bias"], 2, 3), calib))
    a = _ir(sd, p + ".conv6", a, 1, True, calib)
    a = _ir(sd, p + ".conv7", a, 1, True, calib)
    if taps is not None:
        taps["audio"] = a
    return a


@torch.no_grad()
def unet_forward(sd: Dict[str, torch.Tensor], img: torch.Tensor, audio: torch.Tensor, calib: bool = False,
                 taps: Optional[dict] = None) -> torch.Tensor:
    """Model.forward, unet.py:208-226: img (B,6,160,160) in [0,1], audio (B,16,32,32) -> sigmoid output (B,3,160,160)."""
    x1 = _ir(sd, "inc.inconv.0", img, 1, False, calib)
    x2 = _dc(sd, "down1.maxpool_conv.0", x1, 2, calib)
    x3 = _dc(sd, "down2.maxpool_conv.0", x2, 2, calib)
    x4 = _dc(sd, "down3.maxpool_conv.0", x3, 2, calib)
    x5 = _dc(sd, "down4.maxpool_conv.0", x4, 2, calib)
    a = audio_forward(sd, audio, calib, taps)
    f = _dc(sd, "fuse_conv.1", _dc(sd, "fuse_conv.0", torch.cat([x5, a], 1), 1, calib), 1, calib)
    u1 = _up(sd, "up1", f, x4, calib)
    u2 = _up(sd, "up2", u1, x3, calib)
    u3 = _up(sd, "up3", u2, x2, calib)
    u4 = _up(sd, "up4", u3, x1, calib)
    if taps is not None:
        taps.update(x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, fuse=f, u1=u1, u2=u2, u3=u3, u4=u4)
    return torch.sigmoid(F.conv2d(u4, sd["outc.conv.weight"], sd["outc.conv.bias"]))


# ------------------------------------------------------------------------------------------------ synthetic weights / inputs
def synth_inputs(batch: int, seed: int = 3) -> Tuple[torch.Tensor, torch.Tensor, np.ndarray]:
    """-> img (B,6,160,160) exactly as LightReal builds it from seeded 168x168 uint8 crops, HuBERT-like audio windows
    (B,16,32,32) ~ N(0, 0.6) and the crops themselves (B,168,168,3) uint8."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (batch, 21, 21, 3)).astype(np.float32)
    faces = np.kron(base, np.ones((1, 8, 8, 1), np.float32)) + rng.normal(0, 12, (batch, 168, 168, 3))
    faces = np.clip(faces, 0, 255).astype(np.uint8)
    img = torch.stack([lightreal_image(f) for f in faces])
    g = torch.Generator().manual_seed(seed)
    audio = (torch.randn(batch, 16, 32, 32, generator=g) * 0.6)
    return img, audio, faces


def synth_state_dict(seed: int = 0, calib_batch: int = 2) -> Dict[str, torch.Tensor]:
    """Seeded, conditioned weights in the reference checkpoint's key scheme (ultralight.pth = Model(6,'hubert').state_dict()):
    He-normal convolutions, BN gamma ~ U(0.7, 1.3) (halved on residual branches), beta ~ N(0.05, 0.2), running statistics
    calibrated layer by layer on a synthetic batch so that activations stay O(1) and the sigmoid output spans (0, 1)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def bn(p, c, scale=1.0):
        sd[p + ".weight"] = (torch.rand(c, generator=g) * 0.6 + 0.7) * scale
        sd[p + ".bias"] = (torch.randn(c, generator=g) * 0.2 + 0.05) * scale
        sd[p + ".running_mean"] = torch.zeros(c)
        sd[p + ".running_var"] = torch.ones(c)
        sd[p + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)

    def ir(p, inp, oup, res, expand):
        hid = inp * expand
        sd[p + ".conv.0.weight"] = torch.randn(hid, inp, 1, 1, generator=g) * math.sqrt(2.0 / inp)
        bn(p + ".conv.1", hid)
        sd[p + ".conv.3.weight"] = torch.randn(hid, 1, 3, 3, generator=g) * math.sqrt(2.0 / 9)
        bn(p + ".conv.4", hid)
        sd[p + ".conv.6.weight"] = torch.randn(oup, hid, 1, 1, generator=g) * math.sqrt(1.0 / hid)
        bn(p + ".conv.7", oup, 0.5 if res else 1.0)

    irs = {p: (inp, oup, res, e) for p, inp, oup, _s, res, e in ir_list()}
    # state_dict order of the reference module: audio_model (conv1, conv2, conv3, bn3, conv4, conv5, bn5, conv6, conv7), fuse_conv, inc, down*, up*, outc
    for p in ("audio_model.conv1", "audio_model.conv2"):
        ir(p, *irs[p])
    sd["audio_model.conv3.weight"] = torch.randn(CH[3], CH[2], 3, 3, generator=g) * math.sqrt(2.0 / (CH[2] * 9))
    sd["audio_model.conv3.bias"] = torch.randn(CH[3], generator=g) * 0.05
    bn("audio_model.bn3", CH[3])
    ir("audio_model.conv4", *irs["audio_model.conv4"])
    sd["audio_model.conv5.weight"] = torch.randn(CH[4], CH[3], 3, 3, generator=g) * math.sqrt(2.0 / (CH[3] * 9))
    sd["audio_model.conv5.bias"] = torch.randn(CH[4], generator=g) * 0.05
    bn("audio_model.bn5", CH[4])
    for p, _inp, _oup, _s, _res, _e in ir_list():
        if p.startswith("audio_model.conv") and p[-1] in "1245":
            continue
        ir(p, *irs[p])
    sd["outc.conv.weight"] = torch.randn(3, CH[0], 1, 1, generator=g) * math.sqrt(2.0 / CH[0])
    sd["outc.conv.bias"] = torch.randn(3, generator=g) * 0.1
    img, audio, _ = synth_inputs(calib_batch, seed=seed + 77)
    unet_forward(sd, img, audio, calib=True)
    return sd


# ------------------------------------------------------------------------------------------------ LightReal glue
def mirror_index(size: int, index: int) -> int:
    """utils/image.py mirror_index: forward, then backward."""
    turn, res = index // size, index % size
    return res if turn % 2 == 0 else size - res - 1


def lightreal_image(crop_u8: np.ndarray) -> torch.Tensor:
    """One element of LightReal.inference_batch's img_batch (ultralight_avatar.py:148-160) from a (168,168,3) uint8 BGR crop:
    channels 0-2 the 160x160 centre / 255, channels 3-5 the same with cv2.rectangle((5,5,150,145), black, filled) — OpenCV fills
    the rectangle x in [5, 154], y in [5, 149] (both ends inclusive)."""
    real = crop_u8[4:164, 4:164].copy()
    masked = real.copy()
    masked[5:150, 5:155] = 0
    real_t = torch.from_numpy(real.transpose(2, 0, 1).astype(np.float32) / 255.0)
    masked_t = torch.from_numpy(masked.transpose(2, 0, 1).astype(np.float32) / 255.0)
    return torch.cat([real_t, masked_t], 0)


def lightreal_inference_batch(sd, faces: Sequence[np.ndarray], index: int, audiofeat_batch: Sequence[np.ndarray]) -> np.ndarray:
    """LightReal.inference_batch, ultralight_avatar.py:141-169 -> float32 (B,160,160,3) = sigmoid output x 255."""
    B, n = len(audiofeat_batch), len(faces)
    img = torch.stack([lightreal_image(faces[mirror_index(n, index + i)]) for i in range(B)])
    audio = torch.stack([torch.from_numpy(np.asarray(a, np.float32).reshape(16, 32, 32)) for a in audiofeat_batch])
    pred = unet_forward(sd, img, audio)
    return pred.numpy().transpose(0, 2, 3, 1) * 255.0


def lightreal_paste(pred_frame: np.ndarray, frame: np.ndarray, crop_u8: np.ndarray, bbox: Sequence[int], resize=None) -> np.ndarray:
    """LightReal.paste_back_frame, ultralight_avatar.py:171-184.  bbox = (x1, y1, x2, y2).  resize(img, (w, h)): cv2.resize in the
    golden generator; by default the bit-exact restatement oracle.paste_ref.resize_linear_u8."""
    if resize is None:
        from .paste_ref import resize_linear_u8
        resize = lambda img, wh: resize_linear_u8(img, wh[0], wh[1])  # noqa: E731
    x1, y1, x2, y2 = (int(v) for v in bbox)
    out = frame.copy()
    crop = crop_u8.copy()
    crop[4:164, 4:164] = np.asarray(pred_frame).astype(np.uint8)
    out[y1:y2, x1:x2] = resize(crop, (x2 - x1, y2 - y1))
    return out


# ------------------------------------------------------------------------------------------------ HuBERT front end
def wav2vec2_normalize(speech: np.ndarray) -> np.ndarray:
    """Wav2Vec2Processor(speech).input_values for hubert-large-ls960-ft (do_normalize=True): zero mean, unit variance,
    (x - mean) / sqrt(var + 1e-7) in float32 (transformers Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm)."""
    x = np.asarray(speech, np.float32)
    return ((x - x.mean()) / np.sqrt(x.var() + 1e-7)).astype(np.float32)


def expected_frames(n_samples: int) -> int:
    """audio2feature.py:21-25: expected_T = (T - (kernel - stride)) // stride with kernel 400, stride 320."""
    return (n_samples - 80) // 320


def conv_frames(n_samples: int) -> int:
    """Output length of HuBERT's 7-layer conv feature extractor (kernels 10,3,3,3,3,2,2; strides 5,2,2,2,2,2,2; no padding)."""
    t = n_samples
    for k, s in ((10, 5), (3, 2), (3, 2), (3, 2), (3, 2), (2, 2), (2, 2)):
        t = (t - k) // s + 1
    return t


def trim_features(hidden: np.ndarray, n_samples: int) -> np.ndarray:
    """audio2feature.py:50-55: pad with zero rows / cut to expected_T (single-clip case: n_samples < 320000)."""
    T = expected_frames(n_samples)
    assert abs(hidden.shape[0] - T) <= 1
    if hidden.shape[0] < T:
        return np.concatenate([hidden, np.zeros((T - hidden.shape[0], hidden.shape[1]), hidden.dtype)], 0)
    return hidden[:T]


def window_rows(length: int, batch: int, start: float, win=(4, 4), mult: float = 2.0) -> np.ndarray:
    """Row indices BaseASR._feature2chunks gathers (base_asr.py:91-157) as HubertASR.run_step calls it: (batch, 16) int."""
    rows = []
    for i in range(batch):
        center = int((i + start) * mult)
        left, right = int(center - win[0] * mult), int(center + win[1] * mult)
        rows.append([min(max(idx, 0), length - 1) for idx in range(left, right)])
    return np.asarray(rows, np.int64)


def psnr_u8(a: np.ndarray, b: np.ndarray) -> float:
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse)
