"""Debug aid (GPU box): clock64 timeline of the GEMM epilogue warps of CTA 0."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mimo_b200 import lib as L, ops

def main():
    lib = L.load()
    dev = "cuda"
    M, N, K = [int(v) for v in os.environ.get("MNK", "196608,960,320").split(",")]
    res = int(os.environ.get("RES", "0"))
    a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half()
    r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N, device=dev, dtype=torch.half)
    fn = lambda: ops.gemm(a, w, residual=r, out=out)
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"gemm {M}x{N}x{K} res={res}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TFLOP/s, {(M*K+N*K+M*N*(2 if res else 1))*2/ms/1e6:.0f} GB/s")
    tr = torch.zeros(4096, dtype=torch.int64, device=dev)
    lib.mimo_debug_gemm_trace(C.c_void_p(tr.data_ptr()))
    fn(); torch.cuda.synchronize()
    lib.mimo_debug_gemm_trace(None)
    t = tr.cpu()[:2048].view(2, 32, 32)
    t0 = int(t[0, 0, 0])
    for g in range(2):
        print(f"-- group {g}: per tile [start, consts, acc_ready | per chunk: begin, regs, staged, fenced+drained, barrier]")
        for lt in list(range(0, 3)) + list(range(10, 14)):
            row = [int(v) - t0 for v in t[g, lt] if int(v) != 0]
            print(f"tile {lt:2d}: {row}")
    full = tr.cpu()
    mma = full[2048:2560].view(32, 16); prod = full[2560:3072]
    print("-- MMA warp per tile [acc free, k-block operands ready..., issued all]")
    for lt in (10, 11, 12):
        print(f"tile {lt:2d}: {[int(v) - t0 for v in mma[lt] if int(v) != 0]}")
    print("-- producer: stage-free times of k-blocks 50..64:", [int(v) - t0 for v in prod[50:65]])
    per = (int(t[0, 30, 0]) - int(t[0, 10, 0])) / 20
    print(f"steady period per tile: {per:.0f} clk")

if __name__ == "__main__":
    main()
