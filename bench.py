#!/usr/bin/env python
"""bench.py — frames/sec of the MIMO denoising path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our engine (one JSON line on rank 0)
  python bench.py --impl reference [--gpus N] --steps K --warmup W   the oracle port on the host cores

Workload (BASELINE.json configs[1]): 512x512, 24 frames, 20 DDIM steps, CFG 3.5, fp16, synthetic PIL inputs,
seeded random weights of the full architecture (denoising UNet3D 1.31 B params, reference UNet 0.86 B, PoseGuider,
sd-vae-ft-mse-shaped VAE, CLIP ViT-L/14 vision tower) — there are no checkpoints or assets offline.

A "step" is one whole clip: Pose2VideoPipeline's work from CLIP/VAE-encode/pose/reference-UNet through 20 denoising
steps to the batched VAE decode. `value` times sample_tensors() with every input already in HBM; `e2e` times the
public __call__ (PIL in -> CPU video tensor out: PIL pre-processing, pinned H2D, D2H of the clip inside the timed
region). Inputs differ per step only by the noise seed; the UNet touches > 2.6 GB of weights + multi-hundred-MB
activations per forward, far beyond the 126 MB L2, so no explicit L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

WIDTH = HEIGHT = 512
FRAMES = 24
DDIM_STEPS = 20
GUIDANCE = 3.5
METRIC = "frames/sec @ 512x512x24f, 20 DDIM steps"
# algorithmic work per output frame at this config (SURVEY.md §8d / BASELINE.md §2)
TFLOP_PER_FRAME = 49.3

SCHED_KW = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False, steps_offset=1,
                prediction_type="v_prediction", rescale_betas_zero_snr=True, timestep_spacing="trailing")
MOTION_KW = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"],
                 temporal_position_encoding=True, temporal_position_encoding_max_len=32, temporal_attention_dim_div=1)


def synthetic_inputs(frames: int, size: int, seed: int = 0):
    """seeded uint8 reference image, pose frames = black with a coloured blob, white backgrounds."""
    import PIL.Image
    rng = np.random.RandomState(seed)
    ref_img = PIL.Image.fromarray(rng.randint(0, 256, (size, size, 3), dtype=np.uint8))
    poses, bks = [], []
    for i in range(frames):
        a = np.zeros((size, size, 3), np.uint8)
        a[size // 4 + i: size // 2 + i, size // 3: size // 3 + size // 8] = rng.randint(11, 256, 3)
        poses.append(PIL.Image.fromarray(a))
        bks.append(PIL.Image.fromarray(np.full((size, size, 3), 255, np.uint8)))
    return ref_img, poses, bks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1439.7), d.get("hbm_gbs", 6566.4), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# =====================================================================================================
# our engine
# =====================================================================================================
def build_pipeline(device, rank: int = 0, world: int = 1):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    from mimo_b200.host import modules as M
    from mimo_b200.host.pipeline import Pose2VideoPipeline
    from mimo_b200.host.scheduler import DDIMScheduler
    torch.manual_seed(42)  # run_animate.py:46 default seed; weights are the modules' seeded default init
    den = M.UNet3DConditionModel(cross_attention_dim=768, use_inflated_groupnorm=True, use_motion_module=True,
                                 motion_module_mid_block=True, motion_module_type="Vanilla",
                                 motion_module_kwargs=MOTION_KW, unet_use_cross_frame_attention=False,
                                 unet_use_temporal_attention=False)
    ref = M.UNet2DConditionModel(cross_attention_dim=768)
    pg = M.PoseGuider(320, 3, (16, 32, 96, 256))
    vae = M.AutoencoderKL()
    clip = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                                                          num_attention_heads=16, image_size=224, patch_size=14,
                                                          projection_dim=768)).eval()
    pipe = Pose2VideoPipeline(vae=vae, image_encoder=clip, reference_unet=ref, denoising_unet=den, pose_guider=pg,
                              scheduler=DDIMScheduler(**SCHED_KW))
    pipe.to(device, dtype=torch.float16)
    if world > 1:
        pipe.enable_frame_sharding(rank, world)
    return pipe


def run_ours(args):
    import torch.distributed as dist

    from mimo_b200 import ops
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    pipe = build_pipeline(device, rank, world)
    ref_img, poses, bks = synthetic_inputs(FRAMES, WIDTH)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def clip_seed(i):
        return torch.Generator().manual_seed(42 + i)

    # ---- device-resident runs (value) ------------------------------------------------------------------
    host = pipe.preprocess(ref_img, poses, bks, WIDTH, HEIGHT, FRAMES, clip_seed(0), torch.float16)
    dev_in = {k: v.to(device) for k, v in host.items()}
    if args.one_clip:
        t0 = time.perf_counter()
        pipe.sample_tensors(dev_in, DDIM_STEPS, GUIDANCE)
        sync()
        if rank == 0:
            print(json.dumps({"one_clip_s": round(time.perf_counter() - t0, 3), "note": "profiling aid, not a bench value"}))
        return
    for i in range(args.warmup):
        pipe.sample_tensors(dev_in, DDIM_STEPS, GUIDANCE)
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for i in range(args.steps):
        out = pipe.sample_tensors(dev_in, DDIM_STEPS, GUIDANCE)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    launches = ops.launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    pipe._collect_timings()
    phases = dict(pipe.timings)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    ms_per_step = ms / args.steps
    value = FRAMES / (ms_per_step / 1e3)

    # ---- end to end through the public API (PIL in, CPU tensor out) -------------------------------------
    pipe(ref_img, poses, bks, WIDTH, HEIGHT, FRAMES, DDIM_STEPS, GUIDANCE, generator=clip_seed(99))  # warm the host path
    sync()
    k_e2e = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for i in range(k_e2e):
        res = pipe(ref_img, poses, bks, WIDTH, HEIGHT, FRAMES, DDIM_STEPS, GUIDANCE, generator=clip_seed(i))
    sync()
    e2e_s = (time.perf_counter() - t0) / k_e2e
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
    assert res.videos.shape == (1, 3, FRAMES, HEIGHT, WIDTH)

    # ---- per-kernel breakdown + roofline of the dominant kernel (one extra, untimed clip with event brackets) ----
    # The brackets must time kernels, not the host: each forward is preceded by a ~40 ms device-side spin so that the
    # host runs ahead and the ~1 400 launches of the forward sit back to back in the stream when they execute.
    den_eng = pipe.denoising_unet.engine()
    orig_impl = den_eng._forward_impl

    def queued_impl(*a):
        torch.cuda._sleep(80_000_000)
        return orig_impl(*a)

    den_eng._forward_impl = queued_impl
    ops.PROFILE = []
    pipe.sample_tensors(dev_in, DDIM_STEPS, GUIDANCE)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    den_eng._forward_impl = orig_impl
    if args.dump_calls and rank == 0:
        with open(args.dump_calls, "w") as fcsv:
            fcsv.write("idx,name,flops,bytes,ms\n")
            for i, (name, fl, by, a, b) in enumerate(prof):
                fcsv.write(f"{i},{name},{fl:.0f},{by:.0f},{a.elapsed_time(b):.4f}\n")
    agg = {}
    for name, fl, by, a, b in prof:
        d = agg.setdefault(name, [0.0, 0.0, 0.0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += fl
        d[2] += by
        d[3] += 1
    total_ms = sum(d[0] for d in agg.values())
    peak_tf, peak_gbs, peak_src = measured_peaks()
    breakdown = {}
    for name, (t_ms, fl, by, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        breakdown[name] = {"ms": round(t_ms, 2), "share": round(t_ms / total_ms, 4), "launches": cnt,
                           "tflops": round(fl / t_ms / 1e9, 1) if fl else None, "gbs": round(by / t_ms / 1e6, 1)}
    # The dominant kernel is gemm_tcgen05_kernel: the plain GEMMs ("gemm") and the implicit-GEMM 3x3 convolutions
    # ("conv3x3") are the same kernel template in its two addressing modes.
    fam = {"gemm_tcgen05_kernel": ("gemm", "conv3x3"), "attn_spatial": ("attn_spatial",)}
    fam_ms = {k: sum(agg[n][0] for n in names if n in agg) for k, names in fam.items()}
    others = {n: agg[n][0] for n in agg if not any(n in names for names in fam.values())}
    dname = max({**fam_ms, **others}.items(), key=lambda kv: kv[1])[0]
    members = fam.get(dname, (dname,))
    dms = sum(agg[n][0] for n in members if n in agg)
    dfl = sum(agg[n][1] for n in members if n in agg)
    dby = sum(agg[n][2] for n in members if n in agg)
    dcnt = sum(agg[n][3] for n in members if n in agg)
    tensor_bound = dname in ("gemm_tcgen05_kernel", "attn_spatial")
    achieved = dfl / dms / 1e9 if tensor_bound else dby / dms / 1e6
    peak = peak_tf if tensor_bound else peak_gbs
    traffic = None  # DRAM bytes per launch of this kernel from the committed ncu capture (profiles/)
    tpath = Path(__file__).resolve().parent / "profiles" / "ncu_traffic.json"
    if tpath.exists():
        traffic = json.loads(tpath.read_text()).get(dname, {}).get("dram_bytes_per_launch")
    roofline = {"kernel": dname, "bound": "tensor" if tensor_bound else "hbm", "achieved": round(achieved, 1),
                "peak": peak, "unit": "TFLOP/s" if tensor_bound else "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "algorithmic_bytes_per_launch": round(dby / dcnt), "peak_source": peak_src,
                "launches": dcnt, "avg_launch_ms": round(dms / dcnt, 4), "share_of_step": round(dms / total_ms, 4),
                "whole_path_frac_of_tensor_peak": round(TFLOP_PER_FRAME * value / (world * peak_tf), 4)}

    if rank != 0:
        return
    cpu = None if args.no_cpu_baseline else cpu_baseline_sample()
    line = {
        "metric": METRIC, "value": round(value, 4), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 2), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": {"workload": "512x512 x 24-frame animate, 20 DDIM steps, CFG 3.5, fp16 (BASELINE.json configs[1])",
                   "frames": FRAMES, "ddim_steps": DDIM_STEPS, "parallelism": f"frames/{world}",
                   "l2": "inputs/weights per forward >> 126 MB L2; no explicit flush"},
        "e2e": {"value": round(FRAMES / e2e_s, 4), "unit": "frames/s", "h2d_bytes_per_step": pipe.io_bytes["h2d"],
                "d2h_bytes_per_step": pipe.io_bytes["d2h"], "clips_timed": k_e2e},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "phases_ms": {k: round(v, 1) for k, v in phases.items()}, "kernels": breakdown,
    }
    print(json.dumps(line))


# =====================================================================================================
# CPU baseline: the oracle port of the reference's PyTorch graph on the host cores
# =====================================================================================================
_CPU_WEIGHTS: dict = {}


def usable_cores() -> int:
    """Host threads this process may really use: CPU affinity, capped by a cgroup CPU quota if there is one
    (`os.cpu_count()` reports the machine, not the container)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for quota_f, period_f in (("/sys/fs/cgroup/cpu.max", None),
                              ("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us")):
        try:
            if period_f is None:
                q, per = Path(quota_f).read_text().split()
            else:
                q, per = Path(quota_f).read_text().strip(), Path(period_f).read_text().strip()
            if q not in ("max", "-1") and float(per) > 0:
                n = min(n, max(1, int(float(q) / float(per))))
            break
        except (OSError, ValueError):
            continue
    return max(1, n)


def pick_cpu_threads() -> int:
    """The thread count the CPU arm runs with: the fastest of {all usable, 1/2, 1/4} on a one-second probe of the
    UNet's dominant op (a 320-channel 3x3 convolution at the sample's 64x64 resolution). With every hardware thread of
    a 128-thread host on this small a problem the oracle ran 30x slower than with 8 threads - that would be an
    unfairly slow baseline."""
    n = usable_cores()
    x = torch.randn(2, 320, 64, 64)
    w = torch.randn(320, 320, 3, 3)
    cands = sorted({n, max(1, n // 2), max(1, n // 4)}, reverse=True)
    times = {c: float("inf") for c in cands}
    with torch.no_grad():
        for _ in range(2):  # two rounds, best of: the first touches cold pages
            for c in cands:
                torch.set_num_threads(c)
                torch.nn.functional.conv2d(x, w, padding=1)  # warm the primitive cache for this thread count
                t0 = time.perf_counter()
                for _ in range(3):
                    torch.nn.functional.conv2d(x, w, padding=1)
                times[c] = min(times[c], time.perf_counter() - t0)
    best = cands[0]
    for c in cands[1:]:  # prefer more threads unless fewer are clearly (25 %) faster
        if times[c] < 0.75 * times[best]:
            best = c
    return best


def cpu_baseline_sample(seed: int = 0) -> dict:
    """Bounded sample of the same workload: one denoising-UNet forward on 1 of the 24 frames (CFG, 64x64
    latents), one reference-UNet pass, one VAE decode frame and one VAE encode frame at 512x512, fp32, on the thread
    count pick_cpu_threads() finds fastest among all / half / a quarter of the usable host threads;
    extrapolated linearly in frames and steps to frames/s. A reported baseline, not a target."""
    from oracle import torch_oracle as O
    cores = pick_cpu_threads()
    torch.set_num_threads(cores)
    cfg, vcfg = O.UNetConfig(), O.VAEConfig()
    if not _CPU_WEIGHTS:  # seeded random weights, built once per process (20 s of the first sample otherwise)
        _CPU_WEIGHTS.update(den=O.make_denoising_unet_sd(cfg, 1), ref=O.make_reference_unet_sd(cfg, 2),
                            vae=O.make_vae_sd(vcfg, 4))
    sd_den, sd_ref, sd_vae = _CPU_WEIGHTS["den"], _CPU_WEIGHTS["ref"], _CPU_WEIGHTS["vae"]
    g = torch.Generator().manual_seed(seed)
    # bounded sample: 1 of the 24 frames at the full 512x512 resolution (64x64 latents); scaling by x24 frames is
    # linear, which is exact for everything except the 24x24 temporal attention (negligible FLOPs).
    f_s, px = 1, 1
    h = w = HEIGHT // 8
    ehs = torch.cat([torch.zeros(1, 1, 768), torch.randn(1, 1, 768, generator=g)])
    with torch.no_grad():
        t0 = time.perf_counter()
        banks = O.reference_unet_banks(sd_ref, torch.randn(2, 4, h, w, generator=g), ehs, cfg)
        t_ref = (time.perf_counter() - t0) * px
        x = torch.randn(2, 8, f_s, h, w, generator=g)
        pose = torch.randn(2, 320, f_s, h, w, generator=g)
        t0 = time.perf_counter()
        O.denoising_unet(sd_den, x, 499, ehs, pose, banks, cfg, cfg=True)
        t_unet = (time.perf_counter() - t0) * px
        t0 = time.perf_counter()
        O.vae_decode(sd_vae, torch.randn(1, 4, h, w, generator=g), vcfg)
        t_dec = (time.perf_counter() - t0) * px
        t0 = time.perf_counter()
        O.vae_encode_mean(sd_vae, torch.randn(1, 3, HEIGHT, WIDTH, generator=g), vcfg)
        t_enc = (time.perf_counter() - t0) * px
    clip_s = DDIM_STEPS * t_unet * (FRAMES / f_s) + FRAMES * t_dec + 2 * t_enc + t_ref
    return {"value": round(FRAMES / clip_s, 6), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"1 UNet3D forward (CFG) + reference UNet + 1 VAE decode + 1 VAE encode on 1 of 24 frames at 64x64 "
                      f"latents (512x512 px), fp32, {cores} threads; measured per frame: unet {t_unet:.1f}s ref "
                      f"{t_ref:.1f}s dec {t_dec:.1f}s enc {t_enc:.1f}s; extrapolated x{FRAMES} frames x{DDIM_STEPS} steps "
                      f"(animate mode: 2 distinct VAE encodes)",
            "extrapolated_clip_seconds": round(clip_s, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline_sample(seed=i)
        if i >= args.warmup:
            vals.append(r)
    v = statistics.mean(x["value"] for x in vals)
    last = vals[-1]
    last["value"] = round(v, 6)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 6), "unit": "frames/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * FRAMES / v, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "512x512 x 24-frame animate, 20 DDIM steps, CFG 3.5 (BASELINE.json configs[1]); "
                                   "oracle port of the reference's PyTorch graph on the host CPU, bounded sample"},
            "cpu_baseline": last,
            "e2e": {"value": round(v, 6), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dump-calls", default=None, help="write every profiled C-ABI call (name, flops, bytes, ms) as CSV")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU sample (development runs)")
    ap.add_argument("--one-clip", action="store_true",
                    help="run exactly one device-resident clip and exit (for `ncu` launch lists; not a bench value)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps, args.warmup = max(1, min(args.steps, 2)), min(args.warmup, 1)
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device. The engine has no CPU fallback (use --impl reference for the CPU port).")
    args.warmup = max(args.warmup, 3)
    run_ours(args)


if __name__ == "__main__":
    main()
