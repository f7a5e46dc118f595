"""Host side of `e2e` on the CPU (no GPU needed): Pose2VideoPipeline.preprocess() as shipped (frames staged straight into
their tensors, LANCZOS on host threads, sampled-CRC dedupe, noise drawn beside the staging) against the plain functions it
must equal (pil_to_uint8 / _dedupe_images / prepare_latents, what round 1 shipped), for the bench's inputs (already
512 x 512) and for inputs that need the resize (1024 x 1024 frames).   python scripts/host_preprocess_bench.py"""
import sys
import time
from pathlib import Path

import numpy as np
import PIL.Image
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from mimo_b200.host import pipeline as P  # noqa: E402
from mimo_b200.host.scheduler import DDIMScheduler  # noqa: E402


def timeit(f, n=5):
    f()
    f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    pipe = P.Pose2VideoPipeline.__new__(P.Pose2VideoPipeline)
    pipe._clip_image_processor = None
    pipe.scheduler, pipe.vae_scale_factor = DDIMScheduler(**bench.SCHED_KW), 8
    g = torch.Generator().manual_seed(0)
    print(f"host threads usable: {bench.usable_cores()}   (pinned memory is unavailable without a GPU: neither path pins here)")
    for name, src in (("bench inputs, 512x512 (no resize needed)", 512), ("1024x1024 inputs (LANCZOS to 512x512)", 1024)):
        ref, poses, bks = bench.synthetic_inputs(24, src)
        if src != 512:
            rng = np.random.RandomState(1)
            bks = [PIL.Image.fromarray(rng.randint(0, 256, (src, src, 3), dtype=np.uint8)) for _ in range(24)]

        def plain():
            first, inverse = P._dedupe_images(bks)
            return (pipe._clip_pixels(ref), P.pil_to_uint8(ref, 512, 512), P.pil_to_uint8([bks[i] for i in first], 512, 512),
                    P.pil_to_uint8(list(poses), 512, 512), pipe.prepare_latents(1, 4, 512, 512, 24, torch.float16, "cpu", g))

        staged = lambda: pipe.preprocess(ref, poses, bks, 512, 512, 24, g, torch.float16)
        a, b = plain(), staged()
        assert torch.equal(a[3], b["pose_u8"]) and torch.equal(a[2], b["bk_unique_u8"]) and torch.equal(a[1], b["ref_u8"])
        print(f"{name}: plain functions {timeit(plain):7.1f} ms   staged preprocess() {timeit(staged):7.1f} ms")


if __name__ == "__main__":
    main()
