"""Frame-sharded execution check (GPU box, torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/mgpu_check.py
Every rank runs the clip un-sharded (replicated) twice, then frame-sharded over all ranks. GroupNorm statistics are
accumulated with fp32 atomics, so even two identical single-GPU runs differ in the last fp16 ulp and the network
amplifies that; the sharded run must agree with the single-GPU run to within a small multiple of that run-to-run
noise (the same kernels see the same operands: only where K/V live differs)."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
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
    widths = (128, 256, 512, 512)
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
    ok = True
    for F_ in (24, 48):  # one window; three windows with wrap-around
        ref_img, poses, bks = _pil_inputs(F_, 128, seed)
        pipe.enable_frame_sharding(0, 1)
        pipe.denoising_unet.engine().shard = (0, 1, None)
        a = pipe(ref_img, poses, bks, 128, 128, F_, 2, 3.5, generator=torch.manual_seed(42)).videos
        lat_a = pipe.last_latents.clone()
        a2 = pipe(ref_img, poses, bks, 128, 128, F_, 2, 3.5, generator=torch.manual_seed(42)).videos
        noise_l = float((lat_a.float() - pipe.last_latents.float()).abs().max())
        noise_v = float((a - a2).abs().max())
        pipe.enable_frame_sharding(rank, world)
        b = pipe(ref_img, poses, bks, 128, 128, F_, 2, 3.5, generator=torch.manual_seed(42)).videos
        lat_b = pipe.last_latents
        dl = float((lat_a.float() - lat_b.float()).abs().max())
        dv = float((a - b).abs().max())
        print(f"rank {rank}: F={F_} sharded vs single-GPU: max|dlatents|={dl:.3e} max|dvideo|={dv:.3e}   "
              f"(single-GPU run to run: {noise_l:.3e} / {noise_v:.3e})", flush=True)
        ok &= dl <= max(4 * noise_l, 5e-2) and dv <= max(4 * noise_v, 2e-2)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
