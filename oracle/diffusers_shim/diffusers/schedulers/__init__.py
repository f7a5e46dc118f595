from .._pipeline import DDIMScheduler  # noqa: F401


class _Unused:
    def __init__(self, *a, **k):
        raise NotImplementedError("oracle shim: the reference only instantiates DDIMScheduler")


DPMSolverMultistepScheduler = EulerAncestralDiscreteScheduler = EulerDiscreteScheduler = _Unused
LMSDiscreteScheduler = PNDMScheduler = _Unused
