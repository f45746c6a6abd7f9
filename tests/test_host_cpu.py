"""CPU: host-side logic of the product that needs no kernel — scheduler tables vs the oracle DDIM, weight packing vs
the oracle's layouts, LoRA merging, the identity-keyed memo that keeps CUDA-graph addresses stable, synthetic init
(product == oracle, tensor for tensor), sharding helpers, C-ABI argument errors that are raised before any launch."""
import math

import pytest
import torch

from oracle import ops_ref
from oracle.ddim import DDIMOracle


def _sched(**kw):
    from imagdressing_b200.scheduler import DDIMScheduler

    return DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                         clip_sample=False, set_alpha_to_one=False, steps_offset=1, **kw)


@pytest.mark.parametrize("n", [50, 20, 4, 250])
def test_scheduler_tables_match_oracle(n):
    s, o = _sched(), DDIMOracle()
    s.set_timesteps(n)
    o.set_timesteps(n)
    assert torch.equal(s.timesteps, o.timesteps)
    assert torch.equal(s.alphas_cumprod, o.alphas_cumprod)
    ts, coef, blend = s.step_tables("cpu")
    assert ts.dtype == torch.float32 and ts.tolist() == [float(t) for t in o.timesteps]
    for i, t in enumerate(o.timesteps.tolist()):
        a_t, a_p = o.alphas(t)
        assert coef[i].tolist() == pytest.approx([math.sqrt(a_t), math.sqrt(1 - a_t), math.sqrt(a_p), math.sqrt(1 - a_p)],
                                                 rel=1e-6)
    # last step of the reference configuration: alpha_prev = alphas_cumprod[0] (set_alpha_to_one=False), not 1
    assert coef[-1, 2].item() == pytest.approx(math.sqrt(float(o.alphas_cumprod[0])), rel=1e-6)
    # inpaint blend row i = add_noise coefficients at t_{i+1}; the last row leaves the latents untouched
    for i in range(n - 1):
        a_n = float(o.alphas_cumprod[int(o.timesteps[i + 1])])
        assert blend[i].tolist() == pytest.approx([math.sqrt(a_n), math.sqrt(1 - a_n)], rel=1e-6)
    assert blend[-1].tolist() == [1.0, 0.0]
    assert s.step_tables("cpu")[1] is coef  # cached per (device, timesteps): stable addresses for graph replay


def test_scheduler_surface_and_errors():
    s = _sched()
    assert s.order == 1 and s.init_noise_sigma == 1.0 and s.config.steps_offset == 1
    x = torch.randn(2, 4, 8, 8)
    assert s.scale_model_input(x, 5) is x
    with pytest.raises(ValueError):
        s.step(x, 981, x)  # set_timesteps not called
    s.set_timesteps(50)
    with pytest.raises(NotImplementedError):
        s.step(x, 981, x, eta=0.5)
    from imagdressing_b200.scheduler import DDIMScheduler

    with pytest.raises(NotImplementedError):
        DDIMScheduler()  # diffusers default clip_sample=True is not on the reference path
    o = DDIMOracle()
    t = torch.tensor([981, 1])
    n = torch.randn_like(x)
    assert torch.allclose(s.add_noise(x, n, t), o.add_noise(x, n, t))
    s2 = DDIMScheduler.from_config(s.config)
    assert torch.equal(s2.alphas_cumprod, s.alphas_cumprod) and s2.config.set_alpha_to_one is False
    # train.py:403-407 builds the training scheduler with rescale_betas_zero_snr=True: terminal SNR is exactly zero
    z = _sched(rescale_betas_zero_snr=True, timestep_spacing="trailing")
    assert float(z.alphas_cumprod[-1]) == pytest.approx(0.0, abs=1e-7)
    z.set_timesteps(50)
    assert int(z.timesteps[0]) == 999 and int(z.timesteps[-1]) == 19


def test_weight_packing_matches_oracle_layouts():
    from imagdressing_b200 import modeling

    g = torch.Generator().manual_seed(0)
    w = torch.randn(128, 64, 3, 3, generator=g)
    assert torch.equal(modeling.pack_conv3x3(w).float(), ops_ref.conv3x3_pack(w).bfloat16().float())
    # tap-major: column (ky*3+kx)*Cin + ci
    assert modeling.pack_conv3x3(w)[5, (1 * 3 + 2) * 64 + 7].item() == w[5, 7, 1, 2].bfloat16().item()
    wg, bg = torch.randn(2 * 256, 64, generator=g), torch.randn(2 * 256, generator=g)
    pw, pb = modeling.pack_geglu(wg, bg)
    ow, ob = ops_ref.geglu_pack(wg, bg)
    assert torch.equal(pw.float(), ow.bfloat16().float()) and torch.equal(pb, ob)
    # packed row 128*j + i (i < 64) is value row 64*j + i; row 128*j + 64 + i is ITS gate row (256 + 64*j + i)
    assert torch.equal(pw[128 + 3].float(), wg[64 + 3].bfloat16().float())
    assert torch.equal(pw[128 + 64 + 3].float(), wg[256 + 64 + 3].bfloat16().float())


def test_lora_merge_equals_reference_composition():
    """q = to_q(x) + lora_scale * up(down(x))  (adapter/attention_processor.py:453) == x @ (W + s * up @ down)^T."""
    from adapter.attention_processor import LoRALinearLayer
    from imagdressing_b200.processors import _merged

    torch.manual_seed(1)
    lin = torch.nn.Linear(64, 96, bias=False)
    lora = LoRALinearLayer(64, 96, rank=8)
    torch.nn.init.normal_(lora.up.weight, std=0.1)  # the reference zero-inits `up`; make the path visible
    x = torch.randn(5, 64)
    ref = lin(x) + 0.3 * lora.up(lora.down(x))
    assert torch.allclose(x @ _merged(lin, lora, 0.3).T, ref, atol=1e-5)
    assert torch.equal(_merged(lin, None, 0.3), lin.weight.detach().float())
    assert torch.equal(_merged(lin, lora, 0.0), lin.weight.detach().float())


def test_tensor_memo_identity_version_and_stable_storage():
    from imagdressing_b200.processors import TensorMemo

    m = TensorMemo()
    t = torch.zeros(4, 8)
    assert m.get(t) is None
    val = torch.ones(4, 8, dtype=torch.bfloat16)
    assert m.put(t, val, extra=3) is val
    assert m.get(t, extra=3) is val and m.get(t, extra=4) is None  # keyed on the extra (e.g. scale) too
    assert m.get(t.clone(), extra=3) is None  # identity, not equality
    t.add_(1)  # in-place update bumps the version: the memo must miss
    assert m.get(t, extra=3) is None
    m.put(t, val)
    m.clear()
    assert m.get(t) is None
    # ... but the value's storage is offered for reuse, so a captured graph keeps reading the same address
    assert m.reusable((4, 8)) is val and m.reusable((4, 9)) is None and m.reusable((4, 8), torch.float32) is None


def test_synthetic_init_is_identical_for_product_and_oracle():
    from imagdressing_b200 import modeling
    from oracle import unet as ou

    cfg = dict(block_out_channels=(32, 64, 64, 64), cross_attention_dim=64, attention_head_dim=8, norm_num_groups=8)
    p, o = modeling.UNet2DConditionModel(**cfg), ou.UNet2DConditionModel(**cfg)
    modeling.init_synthetic_(p, 7)
    ou.init_synthetic_(o, 7)
    sp, so = p.state_dict(), o.state_dict()
    assert sorted(sp) == sorted(so)  # same names (registration order may differ: the init is keyed by NAME)
    assert all(torch.equal(sp[k], so[k]) for k in sp)
    cp, co = modeling.ControlNetModel(**cfg), ou.ControlNetModel(**cfg)
    modeling.init_synthetic_(cp, 2)
    ou.init_synthetic_(co, 2)
    assert all(torch.equal(v, co.state_dict()[k]) for k, v in cp.state_dict().items())
    # a different seed changes every tensor; the same seed reproduces
    q = modeling.UNet2DConditionModel(**cfg)
    modeling.init_synthetic_(q, 8)
    assert not torch.equal(q.state_dict()["conv_in.weight"], sp["conv_in.weight"])


def test_skip_default_init_gives_the_same_weights_once_initialised():
    from imagdressing_b200 import modeling
    from oracle import unet as ou

    cfg = dict(block_out_channels=(32, 64, 64, 64), cross_attention_dim=64, attention_head_dim=8, norm_num_groups=8)
    saved = torch.nn.Linear.reset_parameters
    with modeling.skip_default_init():
        p, o = modeling.UNet2DConditionModel(**cfg), ou.UNet2DConditionModel(**cfg)
    assert torch.nn.Linear.reset_parameters is saved  # restored
    q = modeling.UNet2DConditionModel(**cfg)
    for m in (p, o, q):
        (ou.init_synthetic_ if m is o else modeling.init_synthetic_)(m, 3)
    ref = q.state_dict()
    assert all(torch.equal(v, ref[k]) for k, v in p.state_dict().items())
    assert all(torch.equal(v, ref[k]) for k, v in o.state_dict().items())
    assert all(torch.isfinite(v).all() for v in p.state_dict().values())


def test_sharding_helpers():
    from imagdressing_b200.parallel import sample_seeds, shard_range

    for total, world in ((64, 8), (10, 4), (3, 8), (1, 1)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))  # contiguous, no overlap, no gap
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
        # seeds derive from the GLOBAL sample index: invariant to the number of ranks
        flat = [x for r in range(world) for x in sample_seeds(42, *spans[r])]
        assert flat == list(sample_seeds(42, 0, total))


def test_abi_argument_errors_are_reported_before_any_launch():
    """The C ABI validates shapes / alignment on the host and returns an error code + message (no exceptions cross
    the boundary); none of these calls reaches a kernel launch, so they run without a GPU."""
    from imagdressing_b200 import _lib

    lib = _lib.load()
    buf = (torch.zeros(64, dtype=torch.float32)).data_ptr()
    rc = lib.imagd_gemm_bf16(buf, 7, buf, 8, buf, 8, 8, 8, 7, None, None)  # K = 7, lda = 7: not multiples of 8
    assert rc != 0 and b"multiples of 8" in lib.imagd_last_error()
    rc = lib.imagd_gemm_bf16(None, 8, buf, 8, buf, 8, 8, 8, 8, None, None)
    assert rc != 0 and b"null pointer" in lib.imagd_last_error()
    rc = lib.imagd_conv3x3_bf16(buf, 48, 1, 8, 8, 48, buf, buf, 64, 64, None, None)  # Cin % 64 != 0
    assert rc != 0 and b"Cin" in lib.imagd_last_error()
    rc = lib.imagd_layernorm_bf16(buf, 8, buf, 8, 4, 4096, None, None, 1e-5, None)  # C > 2048
    assert rc != 0 and b"layernorm" in lib.imagd_last_error()
    with pytest.raises(RuntimeError, match="imagd_gemm_bf16"):
        _lib.check(rc if rc != 0 else -1, "imagd_gemm_bf16")
    assert lib.imagd_version() >= 100
