"""Knock-out timings of the wav2lip256 decoder's ConvT / conv layer shapes (diagnostic tool, not part of the product path).

    python -m livetalking_b200.build --diag          # here (cross-compile lib/libltb200_diag.so with -DLTB_HALO_DIAG)
    python tools/diag_layers.py                      # on the GPU box

For every shape: full kernel, then with one role knocked out (LTB_HALO_DIAG bits: 1 no epilogue global I/O, 2 no epilogue,
4 no MMAs, 8 no A (halo) loads, 16 no B (weight) loads) — what the remaining time is tells which resource bounds the layer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    from livetalking_b200 import _capi
    if not os.environ.get("LTB_DIAG_NORMAL_LIB"):
        _capi.LIB_PATH = os.path.join(os.path.dirname(_capi.LIB_PATH), "libltb200_diag.so")
    from livetalking_b200 import engine
    engine.set_device(0)
    rng = np.random.default_rng(0)
    cases = [  # name, N, H, Cin, Cout, transposed, residual
        ("L50 ConvT 160->64 @128", 16, 128, 160, 64, True, False),
        ("L47 ConvT 320->128 @64", 16, 64, 320, 128, True, False),
        ("L44 ConvT 512->256 @32", 16, 32, 512, 256, True, False),
        ("L41 ConvT 768->384 @16", 16, 16, 768, 384, True, False),
        ("L42 conv 384->384 @32 res", 16, 32, 384, 384, False, True),
        ("L45 conv 256->256 @64 res", 16, 64, 256, 256, False, True),
        ("L39 conv 512->512 @16 res", 16, 16, 512, 512, False, True),
        ("L25 conv 256->256 @16 res", 16, 16, 256, 256, False, True),
        ("L28 conv 512->512 @8 res", 16, 8, 512, 512, False, True),
    ]
    variants = [0, 1, 2, 4, 8, 16, 24, 6, 30]
    sel = os.environ.get("LTB_DIAG_CASES")
    if sel:
        cases = [cases[int(i)] for i in sel.split(",")]
    for name, N, H, cin, cout, tr, res in cases:
        x = (rng.standard_normal((N, H, H, cin)) * 0.5).astype(np.float16)
        w = (rng.standard_normal((cin, cout, 3, 3) if tr else (cout, cin, 3, 3)) * 0.05).astype(np.float32)
        b = np.zeros(cout, np.float32)
        r = x if (res and cin == cout) else None
        flops = 2.0 * N * H * H * 9 * cin * cout
        line = [name]
        base = None
        for v in variants:
            os.environ["LTB_HALO_DIAG"] = str(v)
            _, ms = engine.conv2d_f16(x, w, b, stride=(2, 2) if tr else (1, 1), pad=1, transposed=tr, relu=True, res=r, reps=30)
            if v == 0:
                base = ms
            line.append(f"dbg{v}:{ms * 1000:.1f}us")
        print(" ".join(line), f"| full = {flops / 1e9 / base:.0f} TF/s", flush=True)


if __name__ == "__main__":
    main()
