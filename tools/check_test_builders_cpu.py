"""CPU check for the skip_default_init change of the GPU test builders: the models returned by
tests/test_pipeline_gpu.build / tests/test_unet_gpu.build_pair must have bit-identical state_dicts with and without
torch's default initialisation (every parameter is overwritten by init_synthetic_). Run from the repo root."""
import contextlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import test_pipeline_gpu as tp
import test_unet_gpu as tu
from imagdressing_b200 import modeling


def sds(models):
    return [m.state_dict() for group in models if isinstance(group, tuple) for m in group if m is not None]


def compare(name, fast, slow):
    a, b = sds(fast), sds(slow)
    assert len(a) == len(b)
    n = 0
    for x, y in zip(a, b):
        assert list(x) == list(y)
        for k in x:
            assert torch.equal(x[k], y[k]), (name, k)
            n += 1
    print(f"{name}: {len(a)} models, {n} tensors identical")


dev = torch.device("cpu")
for label, fn in (("pipeline build(controlnet=True)", lambda: tp.build(dev, controlnet=True)),
                  ("unet build_pair(with_ref=True)", lambda: (tu.build_pair(dev, 0, True),)),
                  ("unet build_pair(with_ref=False)", lambda: (tu.build_pair(dev, 1, False),))):
    t = time.time()
    fast = fn()
    tf = time.time() - t
    saved = modeling.skip_default_init
    modeling.skip_default_init = contextlib.nullcontext
    try:
        t = time.time()
        slow = fn()
        tsl = time.time() - t
    finally:
        modeling.skip_default_init = saved
    compare(f"{label} [{tf:.0f} s vs {tsl:.0f} s]", fast, slow)
