"""Packed avatar assets — SURVEY.md §8(f) rank 2: "avatar asset format + GPU-resident loader".

The reference keeps an avatar as a directory of per-frame image files and pickles and decodes every file at start-up
(``load_avatar``: avatars/wav2lip_avatar.py:72-88, avatars/musetalk_avatar.py:69-91; ``read_imgs``: utils/image.py:12-23):

    wav2lip :  full_imgs/<i>.png  face_imgs/<i>.png  coords.pkl            bbox order (y1, y2, x1, x2)
    musetalk:  full_imgs/<i>.png  mask/<i>.png  coords.pkl  mask_coords.pkl  latents.pt      bbox order (x1, y1, x2, y2)

``pack_*`` turns such a directory into ONE file, ``avatar.ltbav``, whose sections are the exact arrays the engine uploads
(decoded uint8 pixels, int32 boxes, the latents in their stored dtype), 4 KiB-aligned so that ``load_packed`` can hand out
``np.memmap`` views: start-up is one sequential read feeding ``cudaMemcpy`` instead of n PNG decodes, and N sessions of a
process share the page cache.  The directory readers below restate the reference loaders and are what the pack is
checked against (tests/test_avatar_pack.py: bit-exact round trip, both kinds, ragged mask sizes, corrupt files).

    python -m livetalking_b200.avatar_pack data/avatars/<id> [--kind wav2lip|musetalk]

File layout (little endian):
    0     8 B   magic  b"LTBAV1\\0\\0"
    8     u32   kind (1 wav2lip, 2 musetalk), u32 n frames, u32 H, u32 W, u32 n_sections, u32 reserved
    32    n_sections x 80 B   { char name[16]; u32 dtype; u32 ndim; u64 shape[4]; u64 offset; u64 nbytes; u64 crc32 }
    4096-aligned sections
"""
from __future__ import annotations

import glob
import os
import pickle
import struct
import sys
import zlib
from typing import Dict, List, Sequence, Tuple

import numpy as np

MAGIC = b"LTBAV1\0\0"
KIND_WAV2LIP, KIND_MUSETALK = 1, 2
ALIGN = 4096
_DTYPES = {0: np.uint8, 1: np.int32, 2: np.float16, 3: np.int64, 4: np.float32}
_DCODE = {np.dtype(v): k for k, v in _DTYPES.items()}
_ENTRY = struct.Struct("<16sII4QQQQ")
_HEAD = struct.Struct("<8sIIIIII")


class AvatarPackError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ directory readers
def _sorted_images(folder: str) -> List[str]:
    """the reference's file order: glob '*.[jpJP][pnPN]*[gG]' sorted by the integer file stem (wav2lip_avatar.py:81-86)"""
    files = glob.glob(os.path.join(folder, "*.[jpJP][pnPN]*[gG]"))
    return sorted(files, key=lambda x: int(os.path.splitext(os.path.basename(x))[0]))


def _read_imgs(paths: Sequence[str]) -> List[np.ndarray]:
    """utils/image.py:12-23 — cv2.imread of every path (BGR uint8)."""
    import cv2
    out = []
    for p in paths:
        img = cv2.imread(p)
        if img is None:
            raise AvatarPackError(f"cannot decode {p}")
        out.append(img)
    return out


def read_wav2lip_dir(avatar_path: str):
    """avatars/wav2lip_avatar.py:72-88: (frame_list_cycle, face_list_cycle, coord_list_cycle)."""
    with open(os.path.join(avatar_path, "coords.pkl"), "rb") as f:
        coords = pickle.load(f)
    frames = _read_imgs(_sorted_images(os.path.join(avatar_path, "full_imgs")))
    faces = _read_imgs(_sorted_images(os.path.join(avatar_path, "face_imgs")))
    return frames, faces, coords


def read_musetalk_dir(avatar_path: str):
    """avatars/musetalk_avatar.py:69-91: (frame_list, mask_list, coord_list, mask_coords_list, input_latent_list)."""
    import torch
    latents = torch.load(os.path.join(avatar_path, "latents.pt"), map_location="cpu")
    with open(os.path.join(avatar_path, "coords.pkl"), "rb") as f:
        coords = pickle.load(f)
    with open(os.path.join(avatar_path, "mask_coords.pkl"), "rb") as f:
        mask_coords = pickle.load(f)
    frames = _read_imgs(_sorted_images(os.path.join(avatar_path, "full_imgs")))
    masks = _read_imgs(_sorted_images(os.path.join(avatar_path, "mask")))
    return frames, masks, coords, mask_coords, latents


# ------------------------------------------------------------------------------------------------ writer
def _stack_same(images: Sequence[np.ndarray], what: str) -> np.ndarray:
    shapes = {im.shape for im in images}
    if len(shapes) != 1:
        raise AvatarPackError(f"{what}: images differ in size {sorted(shapes)[:3]}")
    return np.ascontiguousarray(np.stack(images), np.uint8)


def _write(path: str, kind: int, n: int, H: int, W: int, sections: Dict[str, np.ndarray]) -> str:
    if len(sections) > (ALIGN - _HEAD.size) // _ENTRY.size:
        raise AvatarPackError("too many sections")
    entries, blobs, off = [], [], ALIGN
    for name, arr in sections.items():
        arr = np.ascontiguousarray(arr)
        if arr.dtype not in _DCODE or arr.ndim > 4 or len(name.encode()) > 16:
            raise AvatarPackError(f"section {name}: unsupported dtype/rank {arr.dtype}/{arr.ndim}")
        shape = list(arr.shape) + [0] * (4 - arr.ndim)
        raw = arr.tobytes()
        entries.append(_ENTRY.pack(name.encode(), _DCODE[arr.dtype], arr.ndim, *shape, off, len(raw), zlib.crc32(raw)))
        blobs.append((off, raw))
        off = (off + len(raw) + ALIGN - 1) // ALIGN * ALIGN
    head = _HEAD.pack(MAGIC, kind, n, H, W, len(entries), 0) + b"".join(entries)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(head.ljust(ALIGN, b"\0"))
        for o, raw in blobs:
            f.seek(o)
            f.write(raw)
        f.truncate(off)
    os.replace(tmp, path)
    return path


def pack_wav2lip_lists(frames, faces, coords, out_path: str) -> str:
    fr = _stack_same(frames, "full_imgs")
    fa = _stack_same(faces, "face_imgs")
    co = np.ascontiguousarray(np.asarray(coords), np.int32)
    n = fr.shape[0]
    if fa.shape[0] != n or co.shape != (n, 4):
        raise AvatarPackError(f"frame/face/coords counts differ: {fr.shape[0]}/{fa.shape[0]}/{co.shape}")
    return _write(out_path, KIND_WAV2LIP, n, fr.shape[1], fr.shape[2], {"frames": fr, "faces": fa, "coords": co})


def pack_wav2lip(avatar_path: str, out_path: str = None) -> str:
    frames, faces, coords = read_wav2lip_dir(avatar_path)
    return pack_wav2lip_lists(frames, faces, coords, out_path or os.path.join(avatar_path, "avatar.ltbav"))


def pack_musetalk_lists(frames, masks, coords, mask_coords, latents, out_path: str) -> str:
    fr = _stack_same(frames, "full_imgs")
    n = fr.shape[0]
    co = np.ascontiguousarray(np.asarray(coords), np.int32)
    mc = np.ascontiguousarray(np.asarray(mask_coords), np.int32)
    if len(masks) != n or co.shape != (n, 4) or mc.shape != (n, 4) or len(latents) != n:
        raise AvatarPackError("frame/mask/coords/mask_coords/latents counts differ")
    offs, flat = [0], []
    shapes = np.zeros((n, 3), np.int32)
    for i, m in enumerate(masks):                       # masks are crops of different sizes: concatenated + offset table
        m = np.ascontiguousarray(m, np.uint8)
        shapes[i, :m.ndim] = m.shape
        flat.append(m.reshape(-1))
        offs.append(offs[-1] + m.size)
    lat = np.concatenate([np.asarray(l.detach().cpu().numpy() if hasattr(l, "detach") else l) for l in latents], 0)
    if lat.dtype not in (np.float16, np.float32):
        lat = lat.astype(np.float32)
    return _write(out_path, KIND_MUSETALK, n, fr.shape[1], fr.shape[2],
                  {"frames": fr, "coords": co, "mask_coords": mc, "masks": np.concatenate(flat), "mask_off": np.asarray(offs, np.int64),
                   "mask_shape": shapes, "latents": np.ascontiguousarray(lat)})


def pack_musetalk(avatar_path: str, out_path: str = None) -> str:
    return pack_musetalk_lists(*read_musetalk_dir(avatar_path), out_path or os.path.join(avatar_path, "avatar.ltbav"))


# ------------------------------------------------------------------------------------------------ reader
class PackedAvatar:
    """Header + zero-copy (memory-mapped, read-only) views of the sections."""

    def __init__(self, path: str, verify: bool = False):
        self.path = path
        size = os.path.getsize(path)
        with open(path, "rb") as f:
            head = f.read(ALIGN)
        if len(head) < ALIGN or head[:8] != MAGIC:
            raise AvatarPackError(f"{path}: not an LTBAV1 avatar pack")
        _, self.kind, self.n, self.H, self.W, nsec, _ = _HEAD.unpack_from(head, 0)
        if self.kind not in (KIND_WAV2LIP, KIND_MUSETALK) or nsec > (ALIGN - _HEAD.size) // _ENTRY.size:
            raise AvatarPackError(f"{path}: bad header")
        self.sections: Dict[str, np.ndarray] = {}
        self._where: Dict[str, tuple] = {}
        for i in range(nsec):
            name, dcode, ndim, s0, s1, s2, s3, off, nbytes, crc = _ENTRY.unpack_from(head, _HEAD.size + i * _ENTRY.size)
            name = name.rstrip(b"\0").decode()
            if dcode not in _DTYPES or ndim > 4 or off % ALIGN or off + nbytes > size:
                raise AvatarPackError(f"{path}: section {name} is corrupt or truncated")
            shape = (s0, s1, s2, s3)[:ndim]
            dt = np.dtype(_DTYPES[dcode])
            if int(np.prod(shape, dtype=np.int64)) * dt.itemsize != nbytes:
                raise AvatarPackError(f"{path}: section {name}: shape/size mismatch")
            arr = np.memmap(path, dtype=dt, mode="r", offset=off, shape=shape) if nbytes else np.zeros(shape, dt)
            if verify and zlib.crc32(arr.tobytes()) != crc:
                raise AvatarPackError(f"{path}: section {name}: checksum mismatch")
            self.sections[name] = arr
            self._where[name] = (dt, off, shape, nbytes)
        need = ("frames", "faces", "coords") if self.kind == KIND_WAV2LIP else \
            ("frames", "coords", "mask_coords", "masks", "mask_off", "mask_shape", "latents")
        missing = [k for k in need if k not in self.sections]
        if missing:
            raise AvatarPackError(f"{path}: missing sections {missing}")
        if self.sections["frames"].shape != (self.n, self.H, self.W, 3):
            raise AvatarPackError(f"{path}: frames section does not match the header")

    def host_frames(self) -> list:
        """The ``frame_list_cycle`` the host keeps: WRITABLE per-frame arrays.  On silent frames the reference hands
        ``frame_list_cycle[idx]`` itself to ``cv2.putText`` (avatars/base_avatar.py:417, 449), and OpenCV rejects read-only
        arrays — so the list is a copy-on-write mapping (``mode='c'``: pages are private once written, the file is never
        modified, untouched pages stay shared with the page cache).  The read-only mapping in ``sections`` feeds the upload."""
        dt, off, shape, nbytes = self._where["frames"]
        if not nbytes:
            return []
        return list(np.memmap(self.path, dtype=dt, mode="c", offset=off, shape=shape))

    # the tuples the reference's load_avatar returns (lists of per-frame arrays; views, no copies)
    def wav2lip_lists(self) -> Tuple[list, list, list]:
        if self.kind != KIND_WAV2LIP:
            raise AvatarPackError("not a wav2lip avatar pack")
        s = self.sections
        return self.host_frames(), list(s["faces"]), [tuple(int(v) for v in c) for c in s["coords"]]

    def musetalk_lists(self):
        if self.kind != KIND_MUSETALK:
            raise AvatarPackError("not a musetalk avatar pack")
        s = self.sections
        off, shp = s["mask_off"], s["mask_shape"]
        masks = [s["masks"][off[i]:off[i + 1]].reshape(tuple(int(v) for v in shp[i])) for i in range(self.n)]
        latents = [s["latents"][i:i + 1] for i in range(self.n)]
        return (self.host_frames(), masks, [tuple(int(v) for v in c) for c in s["coords"]],
                [tuple(int(v) for v in c) for c in s["mask_coords"]], latents)


def load_packed(path: str, verify: bool = False) -> PackedAvatar:
    return PackedAvatar(path, verify=verify)


def main(argv: Sequence[str]) -> int:
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0
    kind = "wav2lip"
    if "--kind" in argv:
        kind = argv[argv.index("--kind") + 1]
    path = argv[0]
    out = pack_wav2lip(path) if kind == "wav2lip" else pack_musetalk(path)
    p = load_packed(out, verify=True)
    print(f"{out}: kind={p.kind} n={p.n} {p.W}x{p.H} {os.path.getsize(out) / 1e6:.1f} MB")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
