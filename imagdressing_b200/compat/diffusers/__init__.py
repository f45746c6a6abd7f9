"""Opt-in stand-in for the `diffusers` names the reference's inference scripts import (SURVEY.md §8b "Import surface of
the callers"): put `imagdressing_b200/compat` on PYTHONPATH *instead of* a diffusers install and

    from diffusers import UNet2DConditionModel, AutoencoderKL, DDIMScheduler, ControlNetModel

resolves to the B200 host models of this repo. Only the hot-path classes are real; the VAE (`AutoencoderKL`) is an edge of
the path and must be supplied by the caller — constructing the placeholder explains that.
Never shadow a real diffusers install with this package by accident: it is not on the path unless you add it.
"""
from imagdressing_b200.modeling import ControlNetModel, UNet2DConditionModel  # noqa: F401
from imagdressing_b200.scheduler import DDIMScheduler  # noqa: F401

__version__ = "0.24.0+imagd_b200_compat"


class AutoencoderKL:
    """Placeholder: the VAE is outside the hot path built here. Pass your own module with
    `.encode(x).latent_dist.mean`, `.decode(z, return_dict=False)`, `.config.{scaling_factor, block_out_channels}`,
    or drive the pipelines with `ref_image_latents=` / `output_type="latent"`."""

    def __init__(self, *a, **k):
        raise NotImplementedError(self.__doc__)

    @classmethod
    def from_pretrained(cls, *a, **k):
        raise NotImplementedError(cls.__doc__)


from . import utils, pipelines, schedulers  # noqa: E402,F401
