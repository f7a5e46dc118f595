"""Debug aid (GPU box): time the ping-pong attention variants and dump the clock64 timeline of CTA (0,0,0)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mimo_b200 import lib as L, ops

def timeit(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def main():
    lib = L.load()
    dev = torch.device("cuda:0")
    n, lq, d, heads = 24, 4096, int(os.environ.get("D", 40)), 8
    if d != 40: lq = 1024
    Cc = heads * d
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = (torch.randn(n * lq, 3 * Cc, device=dev, generator=g) * 1.0).half()
    bkv = (torch.randn(1, lq, 2 * Cc, device=dev, generator=g)).half()
    bi = torch.zeros(n, dtype=torch.int32, device=dev)
    out = torch.empty(n * lq, Cc, device=dev, dtype=torch.half)
    fn = lambda: ops.attn_spatial(qkv[:, :Cc], qkv[:, Cc:2*Cc], qkv[:, 2*Cc:], n, lq, heads, bank_k=bkv[:, :, :Cc], bank_v=bkv[:, :, Cc:], bank_index=bi, out=out)
    fl = 4 * n * heads * lq * (2 * lq) * d
    variants = [int(v) for v in os.environ.get("VARIANTS", "0,2,4,6").split(",")]
    for v in variants:
        lib.mimo_debug_attn_variant(v)
        ms = timeit(fn)
        print(f"variant {v}: {ms:.3f} ms = {fl/ms/1e9:.0f} TFLOP/s", flush=True)
    for v in variants:
        lib.mimo_debug_attn_variant(v)
        tr = torch.zeros(5120, dtype=torch.int64, device=dev)
        lib.mimo_debug_attn_trace(C.c_void_p(tr.data_ptr()))
        fn(); torch.cuda.synchronize()
        lib.mimo_debug_attn_trace(None)
        t = tr.cpu()
        sm = t[:4096].view(8, 64, 8)
        mma = t[4096:].view(2, 64, 8)
        t0 = int(sm[0, 0, 0])
        print(f"--- variant {v}: warp A0 (x=0,ew=0) and B0 (x=1,ew=0); columns: wait_s, got_s, s_loaded, exps_done(pre o_wait), o_ok, stored, arrived | per-iteration period")
        for j in list(range(0, 6)) + list(range(30, 36)):
            ra = [int(sm[0, j, k]) - t0 for k in range(7)]
            rb = [int(sm[4, j, k]) - t0 for k in range(7)]
            ma = [int(mma[0, j, k]) - t0 for k in range(6)]
            mb = [int(mma[1, j, k]) - t0 for k in range(6)]
            print(f"j={j:2d} A {ra}  B {rb}\n      issuerA [sfree,kfull,qk,pfull,vfull,pv] {ma}  issuerB {mb}")
        per = (int(sm[0, 60, 0]) - int(sm[0, 10, 0])) / 50
        print(f"steady period per tile (A0): {per:.0f} clk")
    lib.mimo_debug_attn_variant(0)

if __name__ == "__main__":
    main()
