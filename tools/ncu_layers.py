"""Launch ONE wav2lip256 layer shape a few times through the test hook (ltb_conv2d_f16) so that ncu can capture its kernel:

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 1 -o gpurun_out/<name> \
        python tools/ncu_layers.py <case>

cases: gather_s2 (L14 16->32 stride 2 @256, cp.async gather kernel), convt (L50 ConvT 160->64 @128), ystack (L53 80->32 @256),
narrow64 (L52 64->64 + residual @256), wide128 (L49 128->128 + residual @128), s2_tma (L21 64->128 stride 2 @64)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

CASES = {  # N, H, Cin, Cout, k, stride, pad, transposed, residual
    "gather_s2": (16, 256, 16, 32, 3, (2, 2), 1, False, False),
    "convt": (16, 128, 160, 64, 3, (2, 2), 1, True, False),
    "ystack": (16, 256, 80, 32, 3, (1, 1), 1, False, False),
    "narrow64": (16, 256, 64, 64, 3, (1, 1), 1, False, True),
    "wide128": (16, 128, 128, 128, 3, (1, 1), 1, False, True),
    "s2_tma": (16, 64, 64, 128, 3, (2, 2), 1, False, False),          # L21: stride-2 parity-plane TMA mode of the halo kernel
}


def main():
    from livetalking_b200 import engine
    engine.set_device(0)
    N, H, cin, cout, k, s, pad, tr, res = CASES[sys.argv[1]]
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((N, H, H, cin)) * 0.5).astype(np.float16)
    w = (rng.standard_normal((cin, cout, k, k) if tr else (cout, cin, k, k)) * 0.05).astype(np.float32)
    b = np.zeros(cout, np.float32)
    r = x if (res and cin == cout) else None
    _, ms = engine.conv2d_f16(x, w, b, stride=s, pad=pad, transposed=tr, relu=True, res=r, reps=3)
    print(sys.argv[1], f"{ms * 1000:.1f} us per launch")


if __name__ == "__main__":
    main()
