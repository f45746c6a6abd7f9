"""GPU parity of the kernel-backed AutoencoderKL vs the fp32 oracle restatement (oracle/vae.py) at
the real sd-vae-ft-mse widths (128, 256, 512, 512): encode of a 512x512 and a 768x576 image (moments), decode of 64x64 and
96x72 latents (the metric's and configs[3]'s sizes), calibrated against the same oracle run in torch bf16."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def build(dev):
    from imagdressing_b200 import modeling, vae
    from oracle import unet as ou
    from oracle import vae as ov

    with modeling.skip_default_init():
        o, p = ov.AutoencoderKL(), vae.AutoencoderKL()
    ou.init_synthetic_(o, 4)
    p.load_state_dict(o.state_dict())
    return o.to(dev).eval(), p.to(dev).eval()


@pytest.mark.parametrize("hw", [(64, 64), (96, 72)])
@torch.no_grad()
def test_vae_encode_decode(cuda_device, hw):
    dev = cuda_device
    o, p = build(dev)
    h, w = hw
    g = torch.Generator().manual_seed(9)
    img = (torch.rand(1, 3, h * 8, w * 8, generator=g) * 2 - 1).to(dev)
    z = torch.randn(2, 4, h, w, generator=g).to(dev)
    mo = o.encode(img).latent_dist
    mp = p.encode(img).latent_dist
    xo = o.decode(z)[0]
    xp = p.decode(z, return_dict=False)[0]
    ob = o.bfloat16()
    mb = ob.encode(img.bfloat16()).latent_dist
    xb = ob.decode(z.bfloat16())[0]
    e_m, d_m = rel_l2(mp.mean, mo.mean), rel_l2(mb.mean, mo.mean)
    e_x, d_x = rel_l2(xp, xo), rel_l2(xb, xo)
    print(f"{h * 8}x{w * 8}: encode mean rel-L2 {e_m:.4f} (torch-bf16 {d_m:.4f}) | decode rel-L2 {e_x:.4f} (torch-bf16 {d_x:.4f})")
    assert mp.mean.shape == (1, 4, h, w) and xp.shape == (2, 3, h * 8, w * 8)
    assert e_m < max(2 * d_m, 2e-2) and e_x < max(2 * d_x, 2e-2)
    assert rel_l2(mp.logvar, mo.logvar) < max(2 * rel_l2(mb.logvar, mo.logvar), 2e-2)
    assert torch.equal(p.decode(z, return_dict=False)[0], xp)  # deterministic
