"""mimo_exchange (frames <-> pixels re-sharding over peer memory) on the GPU.

One GPU is enough for the kernel itself: G members live in ONE process (their "peer" pointers are ordinary device
pointers), each on its own stream with a small grid so that all members are resident together — the flag protocol,
epoch counting, buffer reuse and the addressing are exactly what runs across GPUs. The model is the numpy
restatement in tests/test_shard_cpu.py. The whole-clip sharded == single-GPU comparison needs >= 2 GPUs
(scripts/mgpu_check.py under torchrun) and is skipped otherwise."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from test_shard_cpu import exchange_model

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _run_group(G, mode, b, fl, hw, C, rounds=3, residual=False, dtype=torch.float16):
    from mimo_b200.host.shard import Exchange
    dev = torch.device("cuda")
    esz = 2
    F_ = fl * G
    rows_src = b * fl * hw if mode != 1 else b * F_ * (hw // G)
    xs = Exchange.local_group(G, {"A": rows_src * C * esz}, dev, timeout_ms=5000, max_blocks=6)
    streams = [torch.cuda.Stream() for _ in range(G)]
    ok = True
    for rnd in range(rounds):  # several exchanges through the same buffers: epochs advance, sources are overwritten
        torch.manual_seed(100 * rnd + G)
        srcs = [torch.randn(rows_src, C, device=dev).to(dtype) for _ in range(G)]
        res = [torch.randn(b * fl * hw, C, device=dev).to(dtype) if residual else None for _ in range(G)]
        rows_dst = {0: b * F_ * (hw // G), 1: b * fl * hw, 2: G * b * fl * hw}[mode]
        dsts = [torch.empty(rows_dst, C, device=dev, dtype=dtype) for _ in range(G)]
        torch.cuda.synchronize()
        for r in range(G):
            with torch.cuda.stream(streams[r]):
                xs[r].bufs["A"].view(rows_src, C, dtype).copy_(srcs[r])
                xs[r].pull(mode, "A", dsts[r], b, fl, hw, C, residual=res[r])
        torch.cuda.synchronize()
        np_srcs = [s.float().cpu().numpy() for s in srcs]
        for r in range(G):
            want = exchange_model(mode, np_srcs, G, r, b, fl, hw, C).reshape(rows_dst, C)
            if residual:
                want = (torch.from_numpy(want).to(dtype).float() + res[r].float().cpu()).to(dtype).float().numpy()
            ok &= np.array_equal(dsts[r].float().cpu().numpy(), want)
    assert int(xs[0].ctl[0]) == rounds + 1  # the device-side epoch advanced once per exchange
    return ok


@pytest.mark.parametrize("G", [2, 4, 8])
def test_exchange_kernel_modes_bit_exact(G):
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    assert _run_group(G, 0, b=1, fl=24 // G, hw=64, C=1280)            # the 8x8 level
    assert _run_group(G, 1, b=1, fl=24 // G, hw=64, C=1280, residual=True)
    assert _run_group(G, 0, b=2, fl=2, hw=1024, C=320)                 # b = 2 (no CFG split), several chunks / segment
    assert _run_group(G, 1, b=2, fl=2, hw=1024, C=320, residual=True)
    assert _run_group(G, 1, b=1, fl=1, hw=8 * G, C=8)                  # tiny, ragged tail of a chunk
    assert _run_group(G, 2, b=1, fl=1, hw=777, C=64)                   # all-gather of the per-step predictions


def test_exchange_single_member_is_a_copy():
    assert _run_group(1, 0, b=2, fl=3, hw=256, C=640)
    assert _run_group(1, 1, b=2, fl=3, hw=256, C=640, residual=True)


def test_exchange_bf16_residual():
    assert _run_group(2, 1, b=1, fl=3, hw=64, C=320, residual=True, dtype=torch.bfloat16)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_clip_equals_single_gpu(world, tmp_path):
    """Whole clips through the public __call__, partitioned over `world` GPUs vs un-sharded, plus determinism of both
    (scripts/mgpu_check.py documents the cases). Needs `world` GPUs on this box."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    out = tmp_path / "mgpu.json"
    port = 29600 + (os.getpid() + world) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "scripts" / "mgpu_check.py"), "--out", str(out),
           "--frames", "24", "48"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, "sharded clip differs from the single-GPU clip (see stdout)"
    res = json.loads(out.read_text())
    assert res["ok"] and all(c["latents_rel_l2"] <= 5e-3 and c["videos_rel_l2"] <= 2e-3 for c in res["cases"])
    keep = os.environ.get("MIMO_MGPU_JSON")
    if keep:
        Path(keep).write_text(json.dumps(res, indent=1))


def test_cfg_branch_split_is_bit_identical():
    """ShardPlan's CFG axis: evaluating one CFG branch alone (batch 1, what a CFG-sharded GPU does) must give exactly the
    rows the two-branch forward gives - every kernel is per row / per image, nothing may depend on the batch."""
    from mimo_b200 import engine as E
    from oracle import torch_oracle as O
    dev = torch.device("cuda")
    widths, f, hw, seed = (128, 256, 512, 512), 24, 16, 41
    cfg = O.UNetConfig(block_out_channels=widths)
    sd_den, sd_ref = O.make_denoising_unet_sd(cfg, seed), O.make_reference_unet_sd(cfg, seed + 1)
    g = torch.Generator().manual_seed(seed)
    ref_lat = torch.randn(1, 4, hw, hw, generator=g).repeat(2, 1, 1, 1).half().to(dev)
    emb = torch.randn(1, 1, 768, generator=g)
    ehs = torch.cat([torch.zeros_like(emb), emb]).half().to(dev)
    x = torch.randn(1, 8, f, hw, hw, generator=g).half().to(dev)
    pose = (torch.randn(f * hw * hw, widths[0], generator=g) * 0.1).half().to(dev)
    den = E.UNetEngine(sd_den, E.UNetSpec(block_out_channels=widths), dev)
    ref = E.UNetEngine(sd_ref, E.UNetSpec(block_out_channels=widths, in_channels=4, motion=False, out_head=False), dev)
    den.use_graphs = False
    banks = ref.write_banks(ref_lat, ehs, den)

    def run(branches):
        den.begin_clip(ehs, banks, cfg=True, frames=f, branches=branches)
        den.taps = {}
        nb = len(branches)
        out = den.forward(x.repeat(nb, 1, 1, 1, 1), 499, pose.repeat(nb, 1).contiguous()).clone()
        taps, den.taps = den.taps, None
        return out, taps

    both, tb = run((0, 1))
    bad = []
    for br in (0, 1):
        one, t1 = run((br,))
        for name in tb:
            if name in t1 and not torch.equal(tb[name][br * f:(br + 1) * f], t1[name]):
                bad.append((br, name, float((tb[name][br * f:(br + 1) * f] - t1[name]).abs().max())))
                break
        if not torch.equal(both[br:br + 1], one):
            bad.append((br, "output", float((both[br:br + 1].float() - one.float()).abs().max())))
    assert not bad, f"first differing block per branch: {bad}"
