"""oracle/ — TEST INFRASTRUCTURE, not product code.

A CPU restatement of the reference's Consistent Self-Attention path (StoryDiffusion's
``SpatialAttnProcessor2_0`` + ``cal_attn_mask_xl`` + ``id_bank``; files cited per function) used ONLY as the
checker by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.  Nothing under ``spider_b200/`` imports it, and the product path fails loudly without its CUDA
library instead of routing through this package.

Pinning: the reference holds no golden vectors or tests for this path.  The restatement is
therefore pinned against the reference ITSELF: ``tests/golden/make_golden.py`` imports the unmodified reference
classes from /root/reference (with gradio/diffusers/spaces/cog stubbed, see ``ref_loader.py``), runs them on CPU
and commits their outputs as fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this package
against those fixtures on every run.
"""
