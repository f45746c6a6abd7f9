"""CPU: nearest-2x upsample + 3x3 conv == four 2x2 phase convs on the low-resolution input with the packed phase weights
(modeling.pack_upconv3x3) under the kernel's phase / tap / offset conventions (tests/emulated_ops.upconv3x3 implements
exactly those) — the algebra and the packing of r2-prep's imagd_upconv3x3_bf16, everything but the CUDA code."""
import pytest
import torch
import torch.nn.functional as F

import emulated_ops


@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(1, 4, 4, 64, 64), (2, 5, 3, 128, 64), (1, 1, 1, 64, 128), (2, 8, 8, 64, 192)])
def test_phase_convs_equal_conv_of_upsampled(NB, H, W, Cin, Cout):
    from imagdressing_b200.modeling import pack_upconv3x3

    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn(NB, H, W, Cin, generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5)
    b = torch.randn(Cout, generator=g)
    up = x.float().permute(0, 3, 1, 2).repeat_interleave(2, 2).repeat_interleave(2, 3)
    ref = F.conv2d(up, w, b, padding=1).permute(0, 2, 3, 1)
    wp = pack_upconv3x3(w)
    assert wp.shape == (4 * Cout, 4 * Cin) and wp.dtype == torch.bfloat16
    out = emulated_ops.upconv3x3(x, wp, bias=b)
    err = float((out.float() - ref).norm() / ref.norm())
    assert err < 6e-3, err  # bf16 rounding of the summed weights and of the output only
    # ... also when the 3x3 weights are bf16 values already (what the model stores)
    wq = w.bfloat16().float()
    sums = pack_upconv3x3(wq).float()
    ref_q = F.conv2d(up, wq, None, padding=1).permute(0, 2, 3, 1)
    y = emulated_ops.upconv3x3(x, sums.bfloat16()).float()
    assert float((y - ref_q).norm() / ref_q.norm()) < 6e-3


def test_upsample_block_uses_phase_path_when_enabled(monkeypatch):
    emulated_ops.install(monkeypatch)
    from imagdressing_b200 import modeling

    torch.manual_seed(0)
    up = modeling.Upsample2D(64)
    x = torch.randn(2, 4, 4, 64).bfloat16()
    plain = up.run(x)
    monkeypatch.setattr(modeling, "UPCONV_PHASE", True)
    up._invalidate()
    phased = up.run(x)
    assert "wp" in up._pk and phased.shape == plain.shape == (2, 8, 8, 64)
    assert float((phased.float() - plain.float()).norm() / plain.float().norm()) < 1e-2
