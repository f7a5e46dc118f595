"""Split-K on/off at the small-M / long-K shapes one GPU of an 8-GPU frame group sees (6 frame-samples): time per call
(CUDA events, L2 flushed). GPU box only."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mimo_b200 import lib as L  # noqa: E402
from mimo_b200 import ops  # noqa: E402
from scripts.gn_bench import timeit  # noqa: E402


def main():
    lib = L.load()
    cases = []
    for n, h, c0, c1, co in [(6, 8, 1280, 0, 1280), (6, 8, 1280, 1280, 1280), (6, 16, 1280, 0, 1280), (6, 16, 1280, 1280, 1280),
                             (6, 16, 1280, 640, 1280), (12, 8, 1280, 0, 1280), (12, 16, 1280, 0, 1280), (6, 32, 640, 0, 640)]:
        x0 = torch.randn(n * h * h, c0, device="cuda").half()
        x1 = torch.randn(n * h * h, c1, device="cuda").half() if c1 else None
        w = torch.randn(co, 9 * (c0 + c1), device="cuda").half()
        b = torch.randn(co, device="cuda").half()
        cases.append((f"conv n={n} {h}x{h} {c0}+{c1}->{co}", 2.0 * n * h * h * co * 9 * (c0 + c1),
                      lambda x0=x0, x1=x1, w=w, b=b, n=n, h=h: ops.conv3x3(x0, w, n, h, h, x1=x1, bias=b)))
    for M, N, K in [(384, 1280, 5120), (1536, 1280, 5120), (384, 1280, 1280), (1536, 1280, 1280), (6144, 640, 2560)]:
        a = torch.randn(M, K, device="cuda").half()
        w = torch.randn(N, K, device="cuda").half()
        b = torch.randn(N, device="cuda").half()
        r = torch.randn(M, N, device="cuda").half()
        cases.append((f"gemm {M}x{N}x{K} +res", 2.0 * M * N * K, lambda a=a, w=w, b=b, r=r: ops.gemm(a, w, bias=b, residual=r)))
    for name, fl, fn in cases:
        row = f"{name:34s}"
        for mode in (0, 1):
            lib.mimo_debug_splitk(mode)
            ms = timeit(fn, iters=10)
            row += f"  split {'on ' if mode else 'off'}: {ms*1e3:8.1f} us {fl/ms/1e9:7.0f} TF/s"
        print(row, flush=True)
    lib.mimo_debug_splitk(1)


if __name__ == "__main__":
    main()
