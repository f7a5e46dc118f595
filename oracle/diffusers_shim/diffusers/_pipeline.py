"""Pipeline-level diffusers symbols for the oracle shim: AutoencoderKL, DDIMScheduler, VaeImageProcessor,
DiffusionPipeline. The arithmetic of the VAE and of DDIM lives once, in oracle/torch_oracle.py; these classes only
give it the object surface the reference's pipeline touches (pipeline_pose2vid_long_edit_bkfill_roiclip.py:71,
373-385, 427-431, 519-521, 551-553, 119-120)."""
from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import PIL.Image
import torch
from torch import nn
from tqdm import tqdm

from ._core import ConfigMixin, FrozenDict, ModelMixin

_ORACLE_DIR = Path(__file__).resolve().parents[2]
if str(_ORACLE_DIR.parent) not in sys.path:
    sys.path.insert(0, str(_ORACLE_DIR.parent))
from oracle import torch_oracle as O  # noqa: E402


class ParamTree(nn.Module):
    """Nested modules whose state_dict() keys are exactly the given dotted names."""

    def __init__(self, sd=None):
        super().__init__()
        for k, v in (sd or {}).items():
            self._put(k.split("."), v)

    def _put(self, parts, v):
        if len(parts) == 1:
            if v.is_floating_point() and parts[0] != "pe":
                self.register_parameter(parts[0], nn.Parameter(v.clone(), requires_grad=False))
            else:
                self.register_buffer(parts[0], v.clone())
            return
        if parts[0] not in self._modules:
            self.add_module(parts[0], ParamTree())
        self._modules[parts[0]]._put(parts[1:], v)


class AutoencoderKL(ModelMixin):
    def __init__(self, state_dict=None, vae_cfg: O.VAEConfig = None, seed: int = 4):
        super().__init__()
        self.vae_cfg = vae_cfg or O.VAEConfig()
        sd = state_dict if state_dict is not None else O.make_vae_sd(self.vae_cfg, seed)
        self.tree = ParamTree(sd)
        self._internal_dict = FrozenDict(block_out_channels=tuple(self.vae_cfg.block_out_channels),
                                         latent_channels=self.vae_cfg.latent_channels, scaling_factor=0.18215)

    @property
    def config(self):
        return self._internal_dict

    def _sd(self):
        return {k[len("tree."):]: v for k, v in self.state_dict().items()}

    def encode(self, x):
        mean = O.vae_encode_mean(self._sd(), x, self.vae_cfg)
        return SimpleNamespace(latent_dist=SimpleNamespace(mean=mean))

    def decode(self, z):
        return SimpleNamespace(sample=O.vae_decode(self._sd(), z, self.vae_cfg))


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 rescale_betas_zero_snr=False, timestep_spacing="leading", **kw):
        assert beta_schedule == "scaled_linear" and prediction_type == "v_prediction" and rescale_betas_zero_snr \
            and timestep_spacing == "trailing" and not clip_sample and set_alpha_to_one, \
            "oracle shim: only the reference's scheduler configuration (inference_v2.yaml:24-33) is restated"
        self._d = O.DDIM(num_train_timesteps, beta_start, beta_end)
        self.alphas_cumprod = self._d.alphas_cumprod
        self.init_noise_sigma = 1.0
        self.timesteps = None
        self.config = FrozenDict(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset)

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(self._d.set_timesteps(num_inference_steps)).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta=0.0, use_clipped_model_output=False, generator=None,
             variance_noise=None, return_dict=True):
        assert eta == 0.0
        prev = self._d.step(model_output, int(timestep), sample)
        return SimpleNamespace(prev_sample=prev)


class VaeImageProcessor(ConfigMixin):
    def __init__(self, do_resize=True, vae_scale_factor=8, resample="lanczos", do_normalize=True,
                 do_binarize=False, do_convert_rgb=False, do_convert_grayscale=False):
        self._internal_dict = FrozenDict(do_resize=do_resize, vae_scale_factor=vae_scale_factor, resample=resample,
                                         do_normalize=do_normalize, do_convert_rgb=do_convert_rgb)

    def preprocess(self, image, height=None, width=None):
        images = image if isinstance(image, list) else [image]
        assert all(isinstance(i, PIL.Image.Image) for i in images)
        c = self.config
        if c.do_convert_rgb:
            images = [i.convert("RGB") for i in images]
        if c.do_resize:
            h = height if height is not None else images[0].height
            w = width if width is not None else images[0].width
            w, h = (x - x % c.vae_scale_factor for x in (w, h))
            images = [i.resize((w, h), resample=PIL.Image.LANCZOS) for i in images]
        arr = np.stack([np.array(i).astype(np.float32) / 255.0 for i in images], axis=0)
        if arr.ndim == 3:
            arr = arr[..., None]
        t = torch.from_numpy(arr.transpose(0, 3, 1, 2))
        if c.do_normalize:
            t = 2.0 * t - 1.0
        return t


class DiffusionPipeline:
    def register_modules(self, **kwargs):
        self._module_names = list(kwargs)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device=None, dtype=None):
        for k in self._module_names:
            m = getattr(self, k)
            if isinstance(m, nn.Module):
                m.to(device=device, dtype=dtype)
        return self

    @property
    def device(self):
        for k in self._module_names:
            m = getattr(self, k)
            if isinstance(m, nn.Module):
                for p in m.parameters():
                    return p.device
        return torch.device("cpu")

    def progress_bar(self, iterable=None, total=None):
        return tqdm(iterable, total=total, disable=True) if iterable is not None else tqdm(total=total, disable=True)
