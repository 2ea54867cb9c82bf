"""SURVEY.md §8(f) rank 2 — packed avatar assets (livetalking_b200/avatar_pack.py) against the reference's directory format.

The "oracle" is the reference loader itself, restated in avatar_pack.read_*_dir (avatars/wav2lip_avatar.py:72-88,
avatars/musetalk_avatar.py:69-91, utils/image.py:12-23): a synthetic avatar is written in the reference's on-disk layout,
read back both ways, and every array must match bit for bit.
"""
import os
import pickle

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from livetalking_b200 import avatar_pack as AP  # noqa: E402


def _write_w2l_dir(root, n=5, H=48, W=64, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "full_imgs"))
    os.makedirs(os.path.join(root, "face_imgs"))
    frames, faces, coords = [], [], []
    for i in range(n):
        fr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        fa = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
        # the reference names files by zero-padded index; sorting is by int(stem): write out of lexical order on purpose
        name = f"{i:08d}.png" if i % 2 == 0 else f"{i}.png"
        cv2.imwrite(os.path.join(root, "full_imgs", name), fr)
        cv2.imwrite(os.path.join(root, "face_imgs", name), fa)
        frames.append(fr)
        faces.append(fa)
        y1, x1 = int(rng.integers(0, 8)), int(rng.integers(0, 8))
        coords.append((y1, y1 + 30, x1, x1 + 40))
    with open(os.path.join(root, "coords.pkl"), "wb") as f:
        pickle.dump(coords, f)
    return frames, faces, coords


def test_wav2lip_pack_round_trip(tmp_path):
    root = str(tmp_path / "avatars" / "w2l")
    os.makedirs(root)
    frames, faces, coords = _write_w2l_dir(root)
    d_frames, d_faces, d_coords = AP.read_wav2lip_dir(root)             # the reference loader, restated
    assert all(np.array_equal(a, b) for a, b in zip(d_frames, frames))   # PNG is lossless: what was written is what it reads
    out = AP.pack_wav2lip(root)
    assert out.endswith("avatar.ltbav") and os.path.getsize(out) % AP.ALIGN == 0
    p = AP.load_packed(out, verify=True)
    assert (p.kind, p.n, p.H, p.W) == (AP.KIND_WAV2LIP, 5, 48, 64)
    fr, fa, co = p.wav2lip_lists()
    assert len(fr) == len(fa) == len(co) == 5
    for i in range(5):
        assert np.array_equal(fr[i], d_frames[i]) and np.array_equal(fa[i], d_faces[i])
        assert tuple(co[i]) == tuple(d_coords[i])                         # (y1, y2, x1, x2) order kept
    # the host's frame list must be WRITABLE: on silent frames the reference draws its watermark straight into
    # frame_list_cycle[idx] (avatars/base_avatar.py:417, 449) and OpenCV rejects read-only arrays.  Copy-on-write mapping:
    # the write is private, the file (and the read-only section that feeds the GPU upload) stays intact.
    assert fr[0].flags.writeable and not fa[0].flags.writeable
    before = fr[0].copy()
    cv2.putText(fr[0], "LiveTalking", (10, 20), cv2.FONT_HERSHEY_SIMPLEX, 0.3, (128, 128, 128), 1)
    assert not np.array_equal(fr[0], before)
    assert np.array_equal(p.sections["frames"][0], before)
    assert np.array_equal(AP.load_packed(out, verify=True).wav2lip_lists()[0][0], before)
    for name, arr in p.sections.items():
        if isinstance(arr, np.memmap):
            assert arr.offset % AP.ALIGN == 0, name


def test_musetalk_pack_round_trip(tmp_path):
    torch = pytest.importorskip("torch")
    root = str(tmp_path / "mt")
    os.makedirs(os.path.join(root, "full_imgs"))
    os.makedirs(os.path.join(root, "mask"))
    rng = np.random.default_rng(3)
    n, H, W = 4, 72, 96
    coords, mask_coords, latents = [], [], []
    for i in range(n):
        cv2.imwrite(os.path.join(root, "full_imgs", f"{i:08d}.png"), rng.integers(0, 256, (H, W, 3), dtype=np.uint8))
        x1, y1 = 20 + i, 10 + i
        x2, y2 = x1 + 30, y1 + 32
        xs, ys, xe, ye = x1 - 5 - i, y1 - 4, x2 + 6, y2 + 3 + i              # crop boxes (and so masks) differ in size per frame
        cv2.imwrite(os.path.join(root, "mask", f"{i:08d}.png"), rng.integers(0, 256, (ye - ys, xe - xs, 3), dtype=np.uint8))
        coords.append((x1, y1, x2, y2))
        mask_coords.append((xs, ys, xe, ye))
        latents.append(torch.from_numpy(rng.standard_normal((1, 8, 32, 32)).astype(np.float16)))
    pickle.dump(coords, open(os.path.join(root, "coords.pkl"), "wb"))
    pickle.dump(mask_coords, open(os.path.join(root, "mask_coords.pkl"), "wb"))
    torch.save(latents, os.path.join(root, "latents.pt"))
    ref = AP.read_musetalk_dir(root)
    p = AP.load_packed(AP.pack_musetalk(root), verify=True)
    assert p.kind == AP.KIND_MUSETALK and p.n == n
    frames, masks, co, mc, lat = p.musetalk_lists()
    for i in range(n):
        assert np.array_equal(frames[i], ref[0][i])
        assert masks[i].shape == ref[1][i].shape and np.array_equal(masks[i], ref[1][i])
        assert tuple(co[i]) == tuple(ref[2][i]) and tuple(mc[i]) == tuple(ref[3][i])     # (x1, y1, x2, y2) / (x_s, y_s, x_e, y_e)
        assert lat[i].shape == (1, 8, 32, 32) and lat[i].dtype == np.float16
        assert np.array_equal(lat[i], ref[4][i].numpy())
    with pytest.raises(AP.AvatarPackError):
        p.wav2lip_lists()


def test_pack_rejects_bad_input(tmp_path):
    root = str(tmp_path / "bad")
    os.makedirs(root)
    frames, faces, coords = _write_w2l_dir(root, n=3)
    out = AP.pack_wav2lip(root)
    raw = open(out, "rb").read()
    bad = str(tmp_path / "magic.ltbav")
    open(bad, "wb").write(b"NOTAVPK1" + raw[8:])
    with pytest.raises(AP.AvatarPackError):
        AP.load_packed(bad)
    trunc = str(tmp_path / "trunc.ltbav")
    open(trunc, "wb").write(raw[: len(raw) // 2])
    with pytest.raises(AP.AvatarPackError):
        AP.load_packed(trunc)
    flipped = bytearray(raw)
    flipped[AP.ALIGN + 100] ^= 0xFF                                       # one corrupted pixel: only the checksum can tell
    corrupt = str(tmp_path / "corrupt.ltbav")
    open(corrupt, "wb").write(bytes(flipped))
    AP.load_packed(corrupt)                                               # structurally fine
    with pytest.raises(AP.AvatarPackError):
        AP.load_packed(corrupt, verify=True)
    with pytest.raises(AP.AvatarPackError):                               # frames of different sizes cannot be stacked
        AP.pack_wav2lip_lists(frames[:2] + [np.zeros((10, 10, 3), np.uint8)], faces, coords, str(tmp_path / "x.ltbav"))
    with pytest.raises(AP.AvatarPackError):                               # count mismatch
        AP.pack_wav2lip_lists(frames, faces[:2], coords, str(tmp_path / "y.ltbav"))


def test_ultralight_directory_packs_with_the_wav2lip_layout(tmp_path):
    """An UltraLight avatar directory (avatars/ultralight_avatar.py:63-82) has the wav2lip layout with 168x168 crops and
    (x1,y1,x2,y2) boxes: the same packer / loader round-trips it bit for bit (plugin.ultralight_avatar.load_avatar prefers the pack)."""
    root = str(tmp_path / "avatars" / "ul")
    rng = np.random.default_rng(3)
    os.makedirs(os.path.join(root, "full_imgs"))
    os.makedirs(os.path.join(root, "face_imgs"))
    frames, faces, coords = [], [], []
    for i in range(4):
        fr, fa = rng.integers(0, 256, (60, 80, 3), dtype=np.uint8), rng.integers(0, 256, (168, 168, 3), dtype=np.uint8)
        cv2.imwrite(os.path.join(root, "full_imgs", f"{i:08d}.png"), fr)
        cv2.imwrite(os.path.join(root, "face_imgs", f"{i:08d}.png"), fa)
        frames.append(fr)
        faces.append(fa)
        coords.append((5 + i, 4, 45 + i, 50))
    with open(os.path.join(root, "coords.pkl"), "wb") as f:
        pickle.dump(coords, f)
    out = AP.pack_wav2lip(root)
    fr2, fa2, co2 = AP.load_packed(out, verify=True).wav2lip_lists()
    assert len(fr2) == len(fa2) == 4 and fa2[0].shape == (168, 168, 3)
    for a, b in zip(frames + faces, list(fr2) + list(fa2)):
        assert np.array_equal(a, b)
    assert [tuple(c) for c in co2] == coords
    assert fr2[0].flags.writeable                       # host frames are copy-on-write (the reference draws into them)
