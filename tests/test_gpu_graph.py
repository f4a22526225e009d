"""One CUDA graph per denoise step (spider_b200/graph.py), on a B200 (`-m gpu`): a captured step replayed K times must
produce exactly what K eager steps produce — same launches, same in-place mask re-sampling from the same generator
stream — and the masks' device buffers must keep their addresses across steps."""
import random

import pytest
import torch

import spider_b200
from spider_b200 import native
from spider_b200.graph import StepGraph
from spider_b200.install import make_processor_class
from oracle.fake_diffusers import FakeAttention

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H = W = 256
FL = 4
LAYERS = [(64, 1280, 20), (256, 640, 10), (64, 1280, 20)]


def _setup(dtype):
    host = spider_b200.StoryGlobals()
    host.height, host.width, host.total_count, host.write = H, W, len(LAYERS), True
    cls = make_processor_class(host)
    torch.manual_seed(3)
    attns = [FakeAttention(c, h).to(DEV, dtype) for (_, c, h) in LAYERS]
    xs = [torch.randn(2 * FL, n, c, device=DEV, dtype=dtype) for (n, c, _) in LAYERS]
    procs = [cls(id_length=FL, device=DEV, dtype=torch.float16) for _ in LAYERS]
    torch.cuda.manual_seed(17)
    host.mask1024, host.mask4096 = spider_b200.cal_attn_mask_xl(FL + 1, FL, 0.5, 0.5, H, W, device=DEV)

    def step():
        host.cur_step, host.attn_count = 25, 0
        return [p(a, x) for p, a, x in zip(procs, attns, xs)]
    return host, procs, step


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_replayed_step_equals_eager_steps(dtype):
    real = random.random
    random.random = lambda: 0.99        # gate open: consistent branch (Comic_Generation.py:98-103)
    try:
        with torch.no_grad():
            host, procs, step = _setup(dtype)
            eager = []
            ptrs0 = [t.data_ptr() for t in host.mask1024.sample_list(DEV)] + [host.mask1024._sample.data_ptr()]
            samples = []
            for _ in range(4):
                samples.append(host.mask4096._sample.clone())
                eager.append([o.clone() for o in step()])
            # the masks were re-sampled in place: same buffers, new contents
            assert [t.data_ptr() for t in host.mask1024.sample_list(DEV)] + [host.mask1024._sample.data_ptr()] == ptrs0
            assert not torch.equal(samples[0], samples[1])
            bank_k = procs[1].id_bank[25].k.clone()
            torch.cuda.synchronize()

            host, procs, step = _setup(dtype)       # same seeds: same weights, inputs, generator stream
            g = StepGraph(step, torch.device(DEV), warmup=1).capture()
            n0 = dict(native.LAUNCHES)
            for r in range(1, 4):                    # the capture's warm-up was step 0
                outs = g.replay()
                torch.cuda.synchronize()
                for li, (got, want) in enumerate(zip(outs, eager[r])):
                    assert torch.equal(got, want), f"replay {r} layer {li}: max diff " \
                        f"{(got.float() - want.float()).abs().max().item():.3e}"
            assert dict(native.LAUNCHES) == n0       # a replay goes through no Python-side launch at all
            assert torch.equal(procs[1].id_bank[25].k, bank_k)
            assert native.debug_stuck() is None
    finally:
        random.random = real
