"""Scheduler names the reference pipelines use in type annotations; only DDIM is implemented (the scripts use DDIM)."""
from imagdressing_b200.scheduler import DDIMScheduler  # noqa: F401


class _NotBuilt:
    def __init__(self, *a, **k):
        raise NotImplementedError("only DDIMScheduler is on the IMAGDressing inference path")


DPMSolverMultistepScheduler = EulerAncestralDiscreteScheduler = EulerDiscreteScheduler = _NotBuilt
LMSDiscreteScheduler = PNDMScheduler = _NotBuilt
