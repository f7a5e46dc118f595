"""Host-side logic of frame sharding on CPU with gloo, world_size 2 (no kernels): the window slices every rank
computes, the rank-major all-gather layout the temporal-attention kernel addresses in place, and the merge of the
per-rank predictions must reproduce the un-sharded ordering."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mimo_b200.host import context


def kv_row(b, f, p, fpc, chunk_stride_rows, hw):
    """Addressing formula of mimo_attn_temporal (include/mimo_b200.h)."""
    return (f // fpc) * chunk_stride_rows + (b * fpc + f % fpc) * hw + p


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    B, hw, C = 2, 3, 4
    for F_ in (24, 48):
        windows = list(context.uniform(0, 20, F_, 24, 1, 4))
        for c in windows:
            fl = len(c) // world
            cl = c[rank * fl:(rank + 1) * fl]
            # "K/V" of this rank: rows ((b, f_local, p)) tagged with (branch, global frame, pixel)
            kv = torch.tensor([[b, cl[f], p, 0] for b in range(B) for f in range(fl) for p in range(hw)], dtype=torch.float32)
            kv_all = torch.empty((world * kv.shape[0], C))
            dist.all_gather_into_tensor(kv_all, kv)
            for b in range(B):
                for f in range(len(c)):
                    for p in range(hw):
                        row = kv_all[kv_row(b, f, p, fl, kv.shape[0], hw)]
                        ok &= row[:3].tolist() == [b, c[f], p]
            # per-window prediction merge: [rep, 4, fl, h, w] per rank -> [rep, 4, f, h, w]
            pred = torch.tensor(cl, dtype=torch.float32).view(1, 1, fl, 1, 1).expand(2, 4, fl, 2, 2).contiguous()
            parts = torch.empty((world * pred.shape[0],) + tuple(pred.shape[1:]))
            dist.all_gather_into_tensor(parts, pred)
            merged = parts.view((world,) + tuple(pred.shape)).permute(1, 2, 0, 3, 4, 5).reshape(2, 4, len(c), 2, 2)
            ok &= merged[0, 0, :, 0, 0].tolist() == [float(x) for x in c]
    results[rank] = ok
    dist.destroy_process_group()


def test_frame_sharding_layout_gloo_world2():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29500 + os.getpid() % 400
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert all(results[r] for r in range(world))


def test_frame_sharding_layout_gloo_world4():
    """Same invariants with four ranks (6 frames of a 24-frame window each): the scaling runs go up to 8 GPUs."""
    world = 4
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29900 + os.getpid() % 90
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert all(results[r] for r in range(world))
