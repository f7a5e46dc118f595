"""Scene compositing of run_edit.py (:253-304) with the per-frame blend chain on the GPU (mimo_composite_frame).

What stays on the host, exactly as in the reference: PIL's resize of the generated frame to the padded clip size, the
crop that removes the padding, the paste on a white canvas at the clip's bounding box, and cv2's INTER_AREA resize of the
16-mode feather mask (tools/util.py:393-437) — small, per-frame, and defined by those libraries' own filters. The blend
(feather mask, occlusion composite, cross-fade of overlapping clips, truncation to uint8) runs as one kernel per frame on
uint8 images that cross PCIe once in each direction."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import ops

MASK_MODES = ("up_down_left_right", "left_right_up", "left_right_down", "up_down_left", "up_down_right", "left_right",
              "up_down", "left_up", "right_up", "left_down", "right_down", "left", "right", "up", "down", "inner")


def mask_mode(bbox: Sequence[int], width: int, height: int) -> int:
    """Index into the 16 feather masks for a clip bounding box (w_min, w_max, h_min, h_max): which image borders the box
    touches (tools/util.py:393-437, same precedence)."""
    w_min, w_max, h_min, h_max = bbox
    L_, R_, U_, D_ = w_min <= 0, w_max >= width, h_min <= 0, h_max >= height
    table = [(L_ and R_ and U_ and D_, 0), (L_ and R_ and U_, 1), (L_ and R_ and D_, 2), (L_ and U_ and D_, 3),
             (R_ and U_ and D_, 4), (L_ and R_, 5), (U_ and D_, 6), (L_ and U_, 7), (R_ and U_, 8), (L_ and D_, 9),
             (R_ and D_, 10), (L_, 11), (R_, 12), (U_, 13), (D_, 14)]
    for hit, idx in table:
        if hit:
            return idx
    return 15


def composite_clip(video: torch.Tensor, context_list: Sequence[Sequence[int]], bbox_clip_list, bk_images_ori,
                   vid_images_ori, occ_mask_images: Optional[list], clip_pad_list, clip_padv_list, mask_list,
                   n_frames: int, overlay: int, device="cuda") -> List[Optional[np.ndarray]]:
    """run_edit.py:253-304. video: [3, n, H, W] float in [0, 1] (the pipeline's output for all clips back to back)."""
    import cv2
    from PIL import Image
    res: List[Optional[torch.Tensor]] = [None] * n_frames
    vi = 0
    for k, context in enumerate(context_list):
        start_i = context[0]
        bbox = bbox_clip_list[k]
        for i in context:
            bk_pil = bk_images_ori[i]
            pad_h, pad_w = clip_pad_list[vi]
            top, bottom, left, right = clip_padv_list[vi]
            image = video[:, vi].permute(1, 2, 0).cpu().numpy()
            frame = Image.fromarray((image * 255).astype(np.uint8)).resize((pad_w, pad_h))
            frame = frame.crop((left, top, pad_w - right, pad_h - bottom))
            w_min, w_max, h_min, h_max = bbox
            canvas = Image.new("RGB", bk_pil.size, "white")
            canvas.paste(frame, (w_min, h_min))
            mask_full = np.zeros((bk_pil.size[1], bk_pil.size[0]), dtype=np.float32)
            mask = cv2.resize(mask_list[mask_mode(bbox, *bk_pil.size)], frame.size, interpolation=cv2.INTER_AREA)
            mask_full[h_min:h_min + mask.shape[0], w_min:w_min + mask.shape[1]] = mask
            to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
            occ = vid = None
            if occ_mask_images is not None:
                occ = to_dev(np.array(occ_mask_images[i])[:, :, 0].astype(np.uint8))
                vid = to_dev(np.array(vid_images_ori[i]))
            factor = (i - start_i + 1) / (overlay + 1)
            res[i] = ops.composite_frame(to_dev(np.array(canvas)), to_dev(np.array(bk_pil)), to_dev(mask_full), occ=occ,
                                         vid=vid, prev=res[i], factor=factor)
            vi += 1
    return [r.cpu().numpy() if r is not None else None for r in res]
