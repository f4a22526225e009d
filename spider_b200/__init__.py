"""spider_b200 — B200-native Consistent Self-Attention (StoryDiffusion path of Layjins/Spider).

Only what the hot path needs: ``csrc/`` (sm_100a CUDA kernels + the C ABI of ``include/csa_b200.h``) and the
host-side mirror of the reference's diffusers attention-processor interface.
"""
__version__ = "0.1.0"
