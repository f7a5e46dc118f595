"""On-disk cache of PACKED engine weights (the reference re-reads and re-lays-out 1.3 B parameters
on every process start, src/models/unet_3d_edit_bkfill.py:578-682).

Packing = fp16/bf16 cast, OIHW -> tap-major [Cout, 9 Cin], q|k|v concatenation, GEGLU tile interleave, parity-class
upsample weights, one [sum(Cout), 1280] time-projection matrix. With MIMO_B200_WEIGHT_CACHE=<dir> the packed tensors of
an engine are written once as a single safetensors file keyed by a fingerprint of the source state dict, and later
starts map that file straight to the device instead of repeating the ~1 300 per-tensor casts, copies and re-layouts.
"""
from __future__ import annotations

import hashlib
import json
import os
from pathlib import Path
from typing import Dict, Optional

import torch


def fingerprint(sd: Dict[str, torch.Tensor], extra: str = "") -> str:
    """Names, shapes, dtypes and EVERY byte of every tensor (a few seconds for the 2.6 GB UNet: a fine-tuned checkpoint
    must never be served another checkpoint's packed weights). MIMO_B200_WEIGHT_CACHE_FASTHASH=1 hashes a strided sample
    (head, tail and 1 024 evenly spaced elements of each tensor) instead - for callers who version their checkpoints."""
    full = os.environ.get("MIMO_B200_WEIGHT_CACHE_FASTHASH") != "1"
    h = hashlib.blake2b(digest_size=16)
    h.update(extra.encode())
    for k in sorted(sd):
        t = sd[k].detach()
        h.update(f"{k}|{tuple(t.shape)}|{t.dtype}".encode())
        flat = t.reshape(-1)
        if flat.numel() == 0:
            continue
        if not full and flat.numel() > 4096:
            idx = torch.linspace(0, flat.numel() - 1, 1024, device=flat.device).long()
            flat = torch.cat([flat[:1024], flat[idx], flat[-1024:]])
        h.update(flat.contiguous().cpu().view(torch.uint8).numpy().tobytes())
    return h.hexdigest()


def _flatten(obj, prefix, out, meta):
    if torch.is_tensor(obj):
        out[prefix] = obj.contiguous()
        return {"t": prefix}
    if obj is None:
        return {"n": None}
    if isinstance(obj, (int, float, str, bool)):
        return {"v": obj}
    if isinstance(obj, dict):
        return {"d": {str(k): _flatten(v, f"{prefix}.{k}" if prefix else str(k), out, meta) for k, v in obj.items()}}
    if isinstance(obj, (list, tuple)):
        return {"l" if isinstance(obj, list) else "u": [_flatten(v, f"{prefix}.{i}", out, meta) for i, v in enumerate(obj)]}
    raise TypeError(f"cannot cache {type(obj)} at {prefix}")


def _unflatten(node, tensors):
    if "t" in node:
        return tensors[node["t"]]
    if "n" in node:
        return None
    if "v" in node:
        return node["v"]
    if "d" in node:
        return {k: _unflatten(v, tensors) for k, v in node["d"].items()}
    seq = [_unflatten(v, tensors) for v in node.get("l", node.get("u"))]
    return seq if "l" in node else tuple(seq)


def cache_dir() -> Optional[Path]:
    d = os.environ.get("MIMO_B200_WEIGHT_CACHE")
    return Path(d) if d else None


def save(path: Path, packed) -> None:
    from safetensors.torch import save_file
    tensors: Dict[str, torch.Tensor] = {}
    tree = _flatten(packed, "", tensors, None)
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_suffix(".tmp")
    save_file({k: v.cpu() for k, v in tensors.items()}, str(tmp), metadata={"tree": json.dumps(tree)})
    os.replace(tmp, path)


def load(path: Path, device):
    from safetensors import safe_open
    with safe_open(str(path), framework="pt", device=str(device)) as f:
        tree = json.loads(f.metadata()["tree"])
        tensors = {k: f.get_tensor(k) for k in f.keys()}
    return _unflatten(tree, tensors)
