from .._core import (SAFETENSORS_WEIGHTS_NAME, USE_PEFT_BACKEND, WEIGHTS_NAME, BaseOutput, deprecate,  # noqa: F401
                     is_accelerate_available, is_torch_version, logging, scale_lora_layers, unscale_lora_layers)
