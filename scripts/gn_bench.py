"""GroupNorm(+SiLU): achieved GB/s (algorithmic 1R + 1W) at the UNet's and the VAE's shapes. GPU box only."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mimo_b200 import lib as L  # noqa: E402
from mimo_b200 import ops  # noqa: E402

SHAPES = [(48, 4096, 320, 0), (48, 4096, 640, 320), (48, 4096, 320, 320), (48, 1024, 640, 0), (48, 1024, 1280, 640),
          (48, 256, 1280, 0), (48, 256, 1280, 1280), (48, 64, 1280, 1280), (6, 4096, 320, 0), (6, 1024, 640, 0),
          (6, 64, 1280, 0), (3, 262144, 128, 0), (24, 65536, 256, 0), (24, 4096, 512, 0)]


def timeit(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def main():
    lib = L.load()
    for n, hw, c0, c1 in SHAPES:
        C = c0 + c1
        x0 = torch.randn(n * hw, c0, device="cuda").half()
        x1 = torch.randn(n * hw, c1, device="cuda").half() if c1 else None
        g, b = torch.randn(C, device="cuda").half(), torch.randn(C, device="cuda").half()
        out = torch.empty(n * hw, C, device="cuda", dtype=torch.half)
        row = f"n={n:3d} hw={hw:6d} C={c0}+{c1}:"
        ms = timeit(lambda: ops.groupnorm(x0, g, b, n, hw, silu=True, x1=x1, out=out))
        row += f"  {ms*1e3:8.1f} us {4.0*n*hw*C/ms/1e6:7.0f} GB/s (algorithmic 1R+1W)"
        print(row, flush=True)


if __name__ == "__main__":
    main()
