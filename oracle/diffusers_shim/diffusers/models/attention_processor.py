from .._core import (ADDED_KV_ATTENTION_PROCESSORS, CROSS_ATTENTION_PROCESSORS, Attention, AttentionProcessor,  # noqa: F401
                     AttnAddedKVProcessor, AttnProcessor, AttnProcessor2_0)
