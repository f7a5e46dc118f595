"""Sharded execution check (GPU box, torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/mgpu_check.py [--out result.json]
Every rank runs clips un-sharded (replicated) twice — which must be bit-identical now that GroupNorm is deterministic —
then partitioned over all ranks (host/shard.py: CFG x windows x frames; frames <-> pixels by mimo_exchange over peer
memory). Cases: CFG on and off, one window (24 frames) and three windows (48 frames, wrap-around), the plan ShardPlan picks
plus the CFG axis forced on. The sharded result may differ from the single-GPU one only by the fp16 rounding of proj_out
before the residual add on the way back from pixels to frames (bounds and their derivation next to the assertion)."""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--full-width", action="store_true", help="full SD1.5 width instead of the reduced test width")
    ap.add_argument("--frames", type=int, nargs="*", default=None, help="only the cases with these frame counts")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import numpy as np
    import PIL.Image

    def _pil_inputs(F_, size, seed):
        rng = np.random.RandomState(seed)
        ref_img = PIL.Image.fromarray(rng.randint(0, 256, (size, size, 3), dtype=np.uint8))
        poses, bks = [], []
        for i in range(F_):
            a = np.zeros((size, size, 3), np.uint8)
            a[size // 4: size // 2 + i % 8, size // 3: size // 3 + 20] = rng.randint(11, 256, 3)
            poses.append(PIL.Image.fromarray(a))
            bks.append(PIL.Image.fromarray(rng.randint(0, 256, (size, size, 3), dtype=np.uint8)))
        return ref_img, poses, bks

    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    from mimo_b200.host import modules as M
    from mimo_b200.host.pipeline import Pose2VideoPipeline
    from mimo_b200.host.scheduler import DDIMScheduler
    from oracle import torch_oracle as O
    widths = (320, 640, 1280, 1280) if args.full_width else (128, 256, 512, 512)
    cfg, vcfg = O.UNetConfig(block_out_channels=widths), O.VAEConfig()
    seed = 700
    mk = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"],
              temporal_position_encoding=True, temporal_position_encoding_max_len=32, temporal_attention_dim_div=1)
    den = M.UNet3DConditionModel(block_out_channels=widths, cross_attention_dim=768, use_inflated_groupnorm=True,
                                 use_motion_module=True, motion_module_mid_block=True, motion_module_type="Vanilla",
                                 motion_module_kwargs=mk)
    ref = M.UNet2DConditionModel(block_out_channels=widths, cross_attention_dim=768)
    pg = M.PoseGuider(widths[0], 3, (16, 32, 96, 256))
    vae = M.AutoencoderKL()
    den.load_state_dict(O.make_denoising_unet_sd(cfg, seed))
    ref.load_state_dict(O.make_reference_unet_sd(cfg, seed + 1))
    pg.load_state_dict(O.make_pose_guider_sd(seed + 2, widths[0]))
    vae.load_state_dict(O.make_vae_sd(vcfg, seed + 3))
    torch.manual_seed(seed + 4)
    clip = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                                          num_attention_heads=4, image_size=224, patch_size=32,
                                                          projection_dim=768)).eval()
    sched = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                          steps_offset=1, prediction_type="v_prediction", rescale_betas_zero_snr=True,
                          timestep_spacing="trailing")
    pipe = Pose2VideoPipeline(vae=vae, image_encoder=clip, reference_unet=ref, denoising_unet=den, pose_guider=pg,
                              scheduler=sched).to(dev, dtype=torch.float16)
    rel = lambda a, b: float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))
    ok, rows = True, []
    size = 128 if world <= 4 else 256  # the coarsest UNet level, (size / 64)^2 tokens, must split over the frame group
    cases = [(24, 3.5, None), (24, 1.0, None), (48, 3.5, None), (48, 1.0, None), (64, 3.5, None)]  # 64: 4 windows
    if world % 2 == 0:  # the CFG axis (chosen last by ShardPlan.make) forced on, with whatever is left on frames
        cases += [(24, 3.5, (2, 1, world // 2)), (48, 3.5, (2, 1, world // 2))]
    if args.frames:
        cases = [c for c in cases if c[0] in args.frames]
    for F_, guidance, forced in cases:
        pipe.force_plan = forced
        ref_img, poses, bks = _pil_inputs(F_, size, seed)
        run = lambda: pipe(ref_img, poses, bks, size, size, F_, 2, guidance, generator=torch.manual_seed(42)).videos
        pipe.enable_sharding(0, 1)
        pipe.force_plan = None
        a = run()
        lat_a = pipe.last_latents.clone()
        a2 = run()
        same = bool(torch.equal(lat_a, pipe.last_latents) and torch.equal(a, a2))
        pipe.enable_sharding(rank, world, exchange_timeout_ms=20000)
        pipe.force_plan = forced
        b1 = run()
        lat_b = pipe.last_latents.clone()
        b2 = run()
        b3 = run()  # eager, capture, replay
        same_sh = bool(torch.equal(lat_b, pipe.last_latents) and torch.equal(b1, b3) and torch.equal(b1, b2))
        from mimo_b200.host.shard import ShardPlan
        from mimo_b200.host.context import uniform
        plan = ShardPlan.make(world, rank, guidance > 1.0, len(list(uniform(0, 2, F_, 24, 1, 4))), 24,
                              min_tokens=(size // 64) ** 2)
        if forced is not None:
            plan = ShardPlan(world, rank, *forced)
        row = {"F": F_, "cfg": guidance > 1.0, "plan": [plan.cfg_ways, plan.win_ways, plan.frame_ways],
               "single_gpu_bit_identical_run_to_run": same, "sharded_bit_identical_run_to_run": same_sh,
               "latents_rel_l2": rel(lat_b, lat_a), "latents_max_abs": float((lat_b.float() - lat_a.float()).abs().max()),
               "videos_rel_l2": rel(b1, a)}
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        # The way back from pixels to frames rounds proj_out's result to fp16 before the residual add (the un-sharded GEMM
        # epilogue adds in fp32): 42 extra roundings per forward out of ~700, measured 1.1e-3 on the latents after two
        # steps; classifier-free guidance (x3.5 on the branch difference) amplifies that to ~3.6e-3. Both are below the
        # distance between two fp16 executions of the same clip (engine vs PyTorch fp16: 2.3e-3 per forward).
        lim = 5e-3 if guidance > 1.0 else 2e-3
        ok &= same and same_sh and row["latents_rel_l2"] <= lim and row["videos_rel_l2"] <= 2e-3
    # every rank must hold the same clip
    chk = torch.tensor([float(pipe.last_latents.float().sum())], device=dev)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok &= float(lo) == float(hi)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and args.out:
        Path(args.out).write_text(json.dumps({"world": world, "ok": bool(flag.item() == 1.0), "cases": rows}, indent=1))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
