class StableDiffusionSafetyChecker:
    """The reference passes this CLASS (not an instance) to the pipeline and never runs it
    (inference_IMAGdressing.py:133-134)."""
