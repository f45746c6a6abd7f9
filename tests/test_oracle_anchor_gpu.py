"""GPU: the anchor of tests/test_oracle_anchor_cpu.py at FULL SD1.5 widths and the metric's 64x64 latent — the
module-tree oracle (oracle/unet.py), the functional restatement (oracle/unet_functional.py) and the PRODUCT kernels are
three implementations of the same forward; all three must agree (the two fp32 oracles to round-off, the bf16 kernel path
within its calibrated tolerance of BOTH)."""
import pytest
import torch

from conftest import rel_l2
from oracle import unet_functional as uf
from test_unet_gpu import build_pair

pytestmark = pytest.mark.gpu


@torch.no_grad()
def test_three_implementations_agree_at_512(cuda_device):
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(1, 4, 64, 64, generator=g).to(dev)
    garment = (torch.randn(1, 4, 64, 64, generator=g) * 0.9).to(dev)
    text = torch.randn(1, 77, 768, generator=g).to(dev)
    gtok = torch.randn(1, 16, 768, generator=g).to(dev)
    t = torch.tensor(481, device=dev)
    ro, rp = build_pair(dev, seed=1, with_ref=False)
    ro(garment, torch.tensor(0, device=dev), gtok)
    rp(garment, torch.tensor(0, device=dev), gtok)
    taps_f = {}
    uf.unet_forward(ro.state_dict(), garment, torch.tensor(0, device=dev), gtok, taps=taps_f)
    names = [n for n in ro.attn_processors if "attn1" in n]
    sa_o = {n: ro.attn_processors[n].cache["hidden_states"] for n in names}
    sa_p = {n: rp.attn_processors[n].cache["hidden_states"] for n in names}
    worst_oo = max(rel_l2(taps_f[n], sa_o[n]) for n in names)
    worst_po = max(rel_l2(sa_p[n], taps_f[n]) for n in names)
    del rp
    o, p = build_pair(dev, seed=0)
    eps_o = o(lat, t, text, cross_attention_kwargs={"sa_hidden_states": sa_o})[0]
    eps_f = uf.unet_forward(o.state_dict(), lat, t, text, garments=sa_o, scale=0.9)
    eps_p = p(lat, t, text, cross_attention_kwargs={"sa_hidden_states": sa_p}, return_dict=False)[0]
    e_oo, e_pf, e_po = rel_l2(eps_f, eps_o), rel_l2(eps_p, eps_f), rel_l2(eps_p, eps_o)
    print(f"taps: functional vs module oracle {worst_oo:.2e}, kernels vs functional {worst_po:.4f} | eps: functional vs "
          f"module oracle {e_oo:.2e}, kernels vs functional {e_pf:.4f}, kernels vs module oracle {e_po:.4f}")
    assert worst_oo < 1e-4 and e_oo < 1e-4
    assert worst_po < 3e-2 and e_pf < 5e-2 and e_po < 5e-2
