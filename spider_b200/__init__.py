"""spider_b200 — B200-native Consistent Self-Attention (StoryDiffusion path of Layjins/Spider).

Only what the hot path needs: ``csrc/`` (sm_100a CUDA kernels + the C ABI of ``include/csa_b200.h``) and the
host-side mirror of the reference's diffusers attention-processor interface.
"""
__version__ = "0.1.0"

from .install import install, make_processor_class, set_attention_processor, uninstall  # noqa: E402,F401
from .graph import StepGraph  # noqa: E402,F401
from .masks import CompactMask, cal_attn_mask_xl  # noqa: E402,F401
from .processor import GLOBALS, SpatialAttnProcessor2_0, StoryGlobals  # noqa: E402,F401
from .stock import AttnProcessor2_0  # noqa: E402,F401
from .lowvram import (CompactIndices, SpatialAttnProcessorLowVram, cal_attn_indice_xl_effcient_memory,  # noqa: E402,F401
                      install_lowvram, load_single_character_weights, make_lowvram_processor_class,
                      save_single_character_weights)
