"""Installing the B200 processor into a StoryDiffusion / SpiderStory pipeline.

The reference instantiates ``SpatialAttnProcessor2_0`` *by name* from the module that also holds the control
globals (StoryDiffusion/Comic_Generation.py:353-371; same in app.py:296-329, predict.py:92-129), so dropping the
new processor in means rebinding that name — and, optionally, ``cal_attn_mask_xl`` so that the first step's masks
are sampled in compact form too — in that host module:

    import StoryDiffusion.Comic_Generation as host
    import spider_b200
    spider_b200.install(host)            # before story_generation(...)
    images = host.story_generation(pipe, general_prompt, prompt_array, style_name)

``set_attention_processor`` mirrors the reference helper of the same name (Comic_Generation.py:270-290) for callers
that install processors themselves.
"""
from __future__ import annotations

import copy
from typing import Optional

from . import masks as _masks
from .processor import GLOBALS, SpatialAttnProcessor2_0
from .stock import AttnProcessor2_0

_CONTROL_DEFAULTS = dict(write=False, cur_step=0, attn_count=0, total_count=0, sa32=0.5, sa64=0.5, height=768,
                         width=768, mask1024=None, mask4096=None)


def make_processor_class(host, bank_store: Optional[str] = None, validate_masks: Optional[bool] = None):
    """A subclass of the B200 processor whose control globals are the attributes of ``host`` (a module or object)."""
    ns = {"_host": host, "__doc__": SpatialAttnProcessor2_0.__doc__, "__module__": SpatialAttnProcessor2_0.__module__}
    if bank_store is not None:
        ns["bank_store"] = bank_store
    if validate_masks is not None:
        ns["validate_masks"] = validate_masks
    return type("SpatialAttnProcessor2_0", (SpatialAttnProcessor2_0,), ns)


def install(host, replace_mask_sampler: bool = True, bank_store: Optional[str] = None,
            validate_masks: Optional[bool] = None, replace_stock_processor: bool = True):
    """Rebind ``host.SpatialAttnProcessor2_0`` (and ``host.cal_attn_mask_xl``, and ``host.AttnProcessor`` — the name
    under which the reference instantiates its stock SDPA processor for every other attention layer,
    Comic_Generation.py:17,368) to the B200 implementations.

    Returns the bound processor class.  Missing control globals are created with the reference's defaults so the
    host namespace is complete even before the driver sets them (:327-349).
    """
    for name, val in _CONTROL_DEFAULTS.items():
        if not hasattr(host, name):
            setattr(host, name, val)
    cls = make_processor_class(host, bank_store, validate_masks)
    if hasattr(host, "SpatialAttnProcessor2_0") and not hasattr(host, "_csa_original_processor"):
        host._csa_original_processor = host.SpatialAttnProcessor2_0
    host.SpatialAttnProcessor2_0 = cls
    if replace_mask_sampler:
        if hasattr(host, "cal_attn_mask_xl") and not hasattr(host, "_csa_original_mask_sampler"):
            host._csa_original_mask_sampler = host.cal_attn_mask_xl
        host.cal_attn_mask_xl = _masks.cal_attn_mask_xl
    if replace_stock_processor:
        if hasattr(host, "AttnProcessor") and not hasattr(host, "_csa_original_stock_processor"):
            host._csa_original_stock_processor = host.AttnProcessor
        host.AttnProcessor = AttnProcessor2_0
    return cls


def uninstall(host) -> None:
    if hasattr(host, "_csa_original_stock_processor"):
        host.AttnProcessor = host._csa_original_stock_processor
        del host._csa_original_stock_processor
    elif getattr(host, "AttnProcessor", None) is AttnProcessor2_0:
        del host.AttnProcessor
    if hasattr(host, "_csa_original_processor"):
        host.SpatialAttnProcessor2_0 = host._csa_original_processor
        del host._csa_original_processor
    if hasattr(host, "_csa_original_mask_sampler"):
        host.cal_attn_mask_xl = host._csa_original_mask_sampler
        del host._csa_original_mask_sampler


def set_attention_processor(unet, id_length: int, host=GLOBALS, all_self_attn: bool = False,
                            other_processor=None, processor_cls=None) -> int:
    """Install processors on ``unet`` the way the reference does (Comic_Generation.py:270-290, :353-371):
    ``up_blocks.*.attn1`` (or every ``attn1`` with ``all_self_attn``) get the consistent-self-attention processor,
    everything else ``other_processor`` — an instance, or ``"b200"`` for a fresh ``spider_b200.AttnProcessor2_0`` per
    layer (the reference's ``AttnProcessor()``, :368, on the same kernels); default: whatever is installed.  Sets
    ``host.total_count`` to the number of consistent processors and returns it."""
    cls = processor_cls
    if cls is None:
        installed = getattr(host, "SpatialAttnProcessor2_0", None)
        if isinstance(installed, type) and issubclass(installed, SpatialAttnProcessor2_0):
            cls = installed                       # the class install(host) bound earlier
        else:
            cls = SpatialAttnProcessor2_0 if host is GLOBALS else make_processor_class(host)
    current = unet.attn_processors
    procs = {}
    count = 0
    for name in current.keys():
        is_self = name.endswith("attn1.processor")
        if is_self and (all_self_attn or name.startswith("up_blocks")):
            procs[name] = cls(id_length=id_length)
            count += 1
        else:
            if isinstance(other_processor, str) and other_processor == "b200":
                procs[name] = AttnProcessor2_0()
            else:
                procs[name] = other_processor if other_processor is not None else current[name]
    unet.set_attn_processor(copy.deepcopy(procs))   # the reference deep-copies the dict (:371)
    host.total_count = count
    return count
