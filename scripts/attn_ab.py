"""A/B of the spatial-attention kernels on the GPU box: variant 32 = one softmax thread per row (attn_spatial_pp.cu),
variant 16 = two threads per row + row sum on the tensor pipe (attn_spatial_pp2.cu), +8 = no start stagger. Correctness against an fp32 torch
reference, then time at the UNet's shapes (CUDA events, L2 flushed between iterations)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mimo_b200 import lib as L  # noqa: E402
from mimo_b200 import ops  # noqa: E402


def ref_attn(q, k, v, n, lq, heads, bank_k=None, bank_v=None, bank_index=None):
    C = q.shape[1]
    d = C // heads
    out = torch.empty(n * lq, C, device=q.device)
    for i in range(n):
        qi = q[i * lq:(i + 1) * lq].float().reshape(lq, heads, d).transpose(0, 1)
        ks = [k[i * lq:(i + 1) * lq].float()]
        vs = [v[i * lq:(i + 1) * lq].float()]
        if bank_k is not None and int(bank_index[i]) >= 0:
            ks.append(bank_k[int(bank_index[i])].float())
            vs.append(bank_v[int(bank_index[i])].float())
        kk = torch.cat(ks).reshape(-1, heads, d).transpose(0, 1)
        vv = torch.cat(vs).reshape(-1, heads, d).transpose(0, 1)
        o = torch.nn.functional.scaled_dot_product_attention(qi[None], kk[None], vv[None])[0]
        out[i * lq:(i + 1) * lq] = o.transpose(0, 1).reshape(lq, C)
    return out


def case(n, lq, lb, heads, d, seed=0, scale_in=1.0):
    torch.manual_seed(seed)
    C = heads * d
    qkv = (torch.randn(n * lq, 3 * C, device="cuda") * scale_in).half()
    bank = (torch.randn(2, lb, 2 * C, device="cuda") * scale_in).half() if lb else None
    bidx = torch.tensor([(-1 if i % 2 == 0 else 1) for i in range(n)], dtype=torch.int32, device="cuda") if lb else None
    return qkv, bank, bidx, C


def run(variant, qkv, bank, bidx, C, n, lq, heads, out=None):
    L.load().mimo_debug_attn_variant(variant)
    kw = {}
    if bank is not None:
        kw = dict(bank_k=bank[:, :, :C], bank_v=bank[:, :, C:], bank_index=bidx)
    return ops.attn_spatial(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], n, lq, heads, out=out, **kw)


def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
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
    ok = True
    for (n, lq, lb, heads, d, sc) in [(2, 256, 256, 8, 40, 1.0), (4, 1024, 1024, 8, 80, 1.0), (2, 4096, 4096, 8, 40, 1.0),
                                      (2, 300, 77, 8, 40, 1.0), (3, 1000, 0, 8, 80, 1.0), (2, 512, 512, 8, 40, 4.0),
                                      (2, 64, 64, 8, 40, 1.0)]:
        qkv, bank, bidx, C = case(n, lq, lb, heads, d, seed=lq + d, scale_in=sc)
        ref = ref_attn(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], n, lq, heads,
                       bank[:, :, :C] if bank is not None else None, bank[:, :, C:] if bank is not None else None, bidx)
        for var in (32, 16):
            got = run(var, qkv, bank, bidx, C, n, lq, heads).float()
            torch.cuda.synchronize()
            e = float((got - ref).norm() / ref.norm())
            bad = not (e < 2e-3)
            ok &= not bad
            print(f"[{'FAIL' if bad else ' ok '}] variant {var:2d} n={n} lq={lq} lb={lb} d={d} x{sc}: rel_l2={e:.3e}", flush=True)
    for (n, lq, lb, heads, d) in [(48, 4096, 4096, 8, 40), (48, 1024, 1024, 8, 80)]:
        qkv, bank, bidx, C = case(n, lq, lb, heads, d, seed=1)
        out = torch.empty(n * lq, C, device="cuda", dtype=torch.half)
        flops = 4.0 * C * lq * (n * lq + (n // 2) * lb)
        for var in (32, 16, 16 | 8):
            ms = timeit(lambda: run(var, qkv, bank, bidx, C, n, lq, heads, out=out))
            print(f"variant {var:2d} n={n} lq={lq} lb={lb} d={d}: {ms:.3f} ms = {flops / ms / 1e9:.0f} TFLOP/s", flush=True)
    L.load().mimo_debug_attn_variant(0)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
