"""Host-side logic of the multi-GPU partitioning on CPU (no kernels): ShardPlan's (CFG x windows x frames) split,
the per-step gather layout, and the addressing of the frames<->pixels exchange (include/mimo_b200.h, mimo_exchange)
restated in numpy. The gloo test (world_size 2 and 4, one process per rank) runs the sampler's gather / scatter with
a stand-in "UNet" whose output tags (window, branch, frame), and must reproduce the serial loop of the reference
(pipeline :492-546) bit for bit."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mimo_b200.host import context
from mimo_b200.host.shard import ShardPlan, gather_layout


def _windows(F_):
    return list(context.uniform(0, 20, F_, 24, 1, 4))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("F_", [24, 48, 64])
@pytest.mark.parametrize("do_cfg", [True, False])
def test_plan_covers_every_window_branch_frame_exactly_once(world, F_, do_cfg):
    windows = _windows(F_)
    plans = [ShardPlan.make(world, r, do_cfg, len(windows), 24) for r in range(world)]
    p0 = plans[0]
    assert p0.cfg_ways * p0.win_ways * p0.frame_ways == world
    assert all((p.cfg_ways, p.win_ways, p.frame_ways) == (p0.cfg_ways, p0.win_ways, p0.frame_ways) for p in plans)
    seen = {}
    for p in plans:
        assert p.rank in p.frame_group() and len(p.frame_group()) == p.frame_ways
        for wi in p.windows_of(len(windows)):
            for br in p.branches(do_cfg):
                for pos, fr in enumerate(p.local_frames(windows[wi])):
                    key = (wi, br, p.frame_idx * (24 // p.frame_ways) + pos)
                    assert key not in seen
                    seen[key] = fr
    assert len(seen) == len(windows) * (2 if do_cfg else 1) * 24
    for (wi, br, pos), fr in seen.items():
        assert fr == windows[wi][pos]  # window order is kept: PE rows are positions inside the window
    # members of a frame group share window and branch and are consecutive ranks
    for p in plans:
        assert {plans[q].coords()[:2] for q in p.frame_group()} == {p.coords()[:2]}


def test_plan_choices_for_the_baseline_configs():
    mk = lambda world, nwin, cfg=True, **kw: (lambda p: (p.cfg_ways, p.win_ways, p.frame_ways))(
        ShardPlan.make(world, 0, cfg, nwin, 24, **kw))
    assert mk(2, 1) == (1, 1, 2)                  # frames before the (unbalanced) CFG pair
    assert mk(4, 1) == (1, 1, 4)
    assert mk(8, 1) == (1, 1, 8)                  # configs[2]: 3 frames x 2 branches per GPU
    assert mk(8, 3) == (1, 1, 8)                  # configs[3]: 48 frames = 3 windows, windows do not divide 8
    assert mk(8, 4) == (1, 4, 2)                  # configs[4]: 64 frames = 4 windows x 2 frame halves
    assert mk(8, 1, False) == (1, 1, 8)
    assert mk(16, 1) == (2, 1, 8)                 # 24 frames do not split 16 ways: the CFG pair takes the last factor
    assert mk(8, 1, min_tokens=4) == (2, 1, 4)    # a 2x2 coarsest level cannot be cut into 8 pixel shards
    with pytest.raises(NotImplementedError):
        ShardPlan.make(5, 0, False, 1, 24)


def exchange_model(mode, srcs, G, r, b, fl, hw, C):
    """numpy restatement of mimo_exchange's three modes for member r (srcs[s] = peer s's source buffer)."""
    hwp, F_ = hw // G, fl * G
    if mode == 0:
        dst = np.zeros((b, F_, hwp, C), srcs[0].dtype)
        for s in range(G):
            dst[:, s * fl:(s + 1) * fl] = srcs[s].reshape(b, fl, hw, C)[:, :, r * hwp:(r + 1) * hwp]
        return dst
    if mode == 1:
        dst = np.zeros((b, fl, hw, C), srcs[0].dtype)
        for s in range(G):
            dst[:, :, s * hwp:(s + 1) * hwp] = srcs[s].reshape(b, F_, hwp, C)[:, r * fl:(r + 1) * fl]
        return dst
    return np.stack([srcs[s].reshape(-1, C) for s in range(G)])


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_exchange_addressing_round_trip(G):
    """frames -> pixels gives every member ALL frames of its pixels, in global frame order; pixels -> frames inverts it."""
    b, fl, hw, C = 2, 24 // G if 24 % G == 0 else 3, 16, 8
    F_ = fl * G
    full = np.arange(b * F_ * hw * C, dtype=np.float32).reshape(b, F_, hw, C)  # the un-sharded token tensor
    local = [full[:, s * fl:(s + 1) * fl].copy() for s in range(G)]              # frame shards
    pix = [exchange_model(0, local, G, r, b, fl, hw, C) for r in range(G)]
    hwp = hw // G
    for r in range(G):
        assert np.array_equal(pix[r], full[:, :, r * hwp:(r + 1) * hwp])
    back = [exchange_model(1, pix, G, r, b, fl, hw, C) for r in range(G)]
    for r in range(G):
        assert np.array_equal(back[r], local[r])


def _serial_reference(F_, do_cfg, h=2, w=2):
    """The reference's accumulation (pipeline :523-546) with a stand-in UNet: pred = f(window, branch, frame)."""
    windows = _windows(F_)
    rep = 2 if do_cfg else 1
    noise = torch.zeros(rep, 4, F_, h, w, dtype=torch.float16)
    counter = torch.zeros(F_, dtype=torch.float16)
    for wi, c in enumerate(windows):
        noise[:, :, c] = noise[:, :, c] + _fake_unet(wi, list(range(rep)), c, h, w)
        counter[c] = counter[c] + 1
    return noise, counter


def _fake_unet(wi, branches, frames, h, w):
    v = torch.tensor([[0.125 * (wi + 1) + 0.5 * br + 0.01 * fr for fr in frames] for br in branches], dtype=torch.float16)
    return v.view(len(branches), 1, len(frames), 1, 1).expand(len(branches), 4, len(frames), h, w).contiguous()


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    h = w = 2
    for F_ in (24, 48, 64):
        for do_cfg in (True, False):
            windows = _windows(F_)
            rep = 2 if do_cfg else 1
            plan = ShardPlan.make(world, rank, do_cfg, len(windows), 24)
            brs = plan.branches(do_cfg)
            mine = plan.windows_of(len(windows))
            stage = torch.stack([_fake_unet(wi, brs, plan.local_frames(windows[wi]), h, w) for wi in mine])
            parts = [torch.empty_like(stage) for _ in range(world)]
            dist.all_gather(parts, stage)  # stands in for mimo_exchange mode 2 over the world group
            gathered = torch.stack(parts)
            noise = torch.zeros(rep, 4, F_, h, w, dtype=torch.float16)
            for q, j, qbrs, frames in gather_layout(plan, windows, do_cfg):
                idx = torch.tensor(frames)
                noise[qbrs[0]:qbrs[-1] + 1] = noise[qbrs[0]:qbrs[-1] + 1].index_add(2, idx, gathered[q, j])
            want, _ = _serial_reference(F_, do_cfg, h, w)
            ok &= torch.equal(noise, want)
    results[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gather_scatter_matches_the_serial_loop_gloo(world):
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29500 + (os.getpid() * 7 + world) % 400
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert all(results[r] for r in range(world))
