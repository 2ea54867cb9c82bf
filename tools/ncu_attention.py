"""Launch the fused attention kernel on one MuseTalk shape a few times so that ncu can capture it:

    ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 1 -c 1 -o gpurun_out/<name> python tools/ncu_attention.py <case>

cases: self1024 (UNet 32x32 self-attention, batch 8 x 8 heads, d 40 padded to 48), whisper (1500 tokens, 6 heads, d 64),
cross (1024 queries x 50 audio keys)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

CASES = {"self1024": (8, 8, 48, 1024, 1024, 1024, 40), "whisper": (1, 6, 64, 1500, 1500, 1500, 64), "cross": (8, 8, 48, 1024, 64, 50, 40)}


def main():
    from livetalking_b200 import engine
    from livetalking_b200.ops import Ctx
    engine.set_device(0)
    B, H, d, nq, kv_rows, valid, d_true = CASES[sys.argv[1] if len(sys.argv) > 1 else "self1024"]
    rng = np.random.default_rng(0)
    ctx = Ctx()
    Hd = H * d
    q = ctx.upload((rng.standard_normal((B, nq, Hd)) * 1.2).astype(np.float16))
    kv = ctx.upload((rng.standard_normal((B, kv_rows, 2 * Hd)) * 1.2).astype(np.float16))
    n_pad = (kv_rows + 15) // 16 * 16
    vt = ctx.alloc((B * H, d, n_pad), np.float16, zero=True)
    ctx.transpose_heads(kv.ptr + 2 * Hd, B, kv_rows, 2 * Hd, H, d, n_pad, vt)
    out = ctx.alloc((B * nq, Hd), np.float16, zero=True)
    for _ in range(4):
        ctx.attention(q.ptr, Hd, kv.ptr, 2 * Hd, kv_rows, vt, n_pad, B, H, nq, valid, d, float(d_true) ** -0.5, out)
    ctx.sync()
    print("ok", float(np.abs(ctx.download(out).astype(np.float32)).mean()))


if __name__ == "__main__":
    main()
