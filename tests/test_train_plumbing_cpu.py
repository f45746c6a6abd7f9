"""CPU: the callers' side of the training step — the IGPair record / batch format
(/root/reference/IGPair.py:12-127), the frozen-encoder batch preparation (train.py:519-560), and DeepSpeed-layout checkpoints
(train.py:179-207; inference_IMAGdressing.py:97-117 reads `["module"]` with `unet.` / `ref_unet.` / `proj.` /
`adapter_modules.` prefixes)."""
import json
import os
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import emulated_ops

BF = torch.bfloat16


class ToyTokenizer:
    model_max_length = 12

    def __call__(self, text, max_length, padding, truncation, return_tensors):
        ids = [1] + [3 + (ord(c) % 50) for c in text][: max_length - 2] + [2]
        ids = ids + [0] * (max_length - len(ids))
        return SimpleNamespace(input_ids=torch.tensor([ids], dtype=torch.long))


def make_records(tmp_path, n=3):
    from PIL import Image

    recs = []
    rng = np.random.default_rng(0)
    for i in range(n):
        paths = []
        for kind, (w, h) in (("person", (600, 800)), ("cloth", (768, 1024))):
            p = tmp_path / f"{kind}_{i}.png"
            Image.fromarray(rng.integers(0, 255, (h, w, 3), dtype=np.uint8)).save(p)
            paths.append(str(p))
        recs.append({"image_file": paths[0], "cloth_file": paths[1], "text": [f"a shirt {i}", f"garment {i}"]})
    f = tmp_path / "data.json"
    f.write_text(json.dumps(recs))
    return str(f)


def test_dataset_items_and_collate(tmp_path):
    from imagdressing_b200.data import VDDataset, collate_fn

    jf = make_records(tmp_path)
    ds = VDDataset([jf, jf], ToyTokenizer(), rng=random.Random(0))
    assert len(ds) == 6
    items = [ds[i] for i in range(6)]
    for it in items:
        assert it["vae_person"].shape == (3, 640, 512) and it["vae_clothes"].shape == (3, 640, 512)
        assert -1.0 <= float(it["vae_person"].min()) and float(it["vae_person"].max()) <= 1.0
        assert it["clip_image"].shape == (1, 3, 224, 224)
        assert it["text_input_ids"].shape == (1, 12) and it["null_text_input_ids"].tolist() == [[1, 2] + [0] * 10]
        assert it["drop_image_embed"] in (0, 1)
    b = collate_fn(items[:4])
    assert b["vae_person"].shape == (4, 3, 640, 512) and b["vae_person"].dtype == torch.float32
    assert b["clip_image"].shape == (4, 3, 224, 224) and b["input_ids"].shape == (4, 12) and b["null_input_ids"].shape == (4, 12)
    assert len(b["drop_image_embed"]) == 4 and len(b["text"]) == 4
    # conditioning dropout rates (IGPair.py:60-68): ~5 % image only, ~5 % text only, ~5 % both
    ds2 = VDDataset(jf, ToyTokenizer(), rng=random.Random(1))
    drops = texts = 0
    n = 400
    for i in range(n):
        r = ds2.rng.random()
        drops += r < 0.05 or 0.1 <= r < 0.15
        texts += 0.05 <= r < 0.15
    assert 0.05 < drops / n < 0.16 and 0.05 < texts / n < 0.16
    with pytest.raises(ValueError):
        VDDataset(123, ToyTokenizer())


def test_prepare_batch_with_stand_in_encoders():
    from imagdressing_b200 import train
    from imagdressing_b200.scheduler import DDIMScheduler

    class Vae(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1))

        def encode(self, x):
            z = torch.nn.functional.avg_pool2d(x, 8)[:, :1].repeat(1, 4, 1, 1)
            return SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda: z))

    class Img(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1))

        def forward(self, x, output_hidden_states=True):
            h = x.mean((2, 3))[:, None, :1].expand(-1, 257, 1280)
            return SimpleNamespace(hidden_states=[h * 0, h, h * 2])

    class Txt(torch.nn.Module):
        def forward(self, ids):
            return (ids[..., None].float().expand(-1, -1, 768),)

    B = 3
    g = torch.Generator().manual_seed(0)
    batch = dict(vae_person=torch.rand(B, 3, 64, 48, generator=g) * 2 - 1, vae_clothes=torch.rand(B, 3, 64, 48, generator=g) * 2 - 1,
                 clip_image=torch.rand(B, 3, 224, 224, generator=g) + 1.0, drop_image_embed=[0, 1, 0],
                 input_ids=torch.randint(0, 50, (B, 12), generator=g))
    sched = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False)
    out = train.prepare_batch(batch, Vae(), Img(), Txt(), sched, torch.device("cpu"), noise_offset=0.05,
                              generator=torch.Generator().manual_seed(1))
    assert out["latents"].shape == (B, 4, 8, 6) and out["ref_latents"].shape == (B, 4, 8, 6) and out["noise"].shape == (B, 4, 8, 6)
    assert out["timesteps"].dtype == torch.long and int(out["timesteps"].max()) < 1000
    assert out["clip_image_embeddings"].shape == (B, 257, 1280) and out["encoder_hidden_states"].shape == (B, 12, 768)
    assert float(out["clip_image_embeddings"][1].abs().max()) == 0.0 and float(out["clip_image_embeddings"][0].abs().max()) > 0  # dropped
    want = torch.nn.functional.avg_pool2d(batch["vae_person"], 8)[:, :1] * 0.18215
    assert torch.allclose(out["latents"][:, :1], want)


def test_checkpoint_round_trip_in_deepspeed_layout(tmp_path, monkeypatch):
    emulated_ops.install(monkeypatch)
    from test_train_cpu import batch, build_pair
    from imagdressing_b200 import train
    from imagdressing_b200.scheduler import DDIMScheduler

    _, (p_unet, p_ref, p_proj, p_ad) = build_pair()
    sd = train.SDModel(p_unet, p_ref, p_proj, p_ad)
    for m in (p_unet, p_ref, p_proj):
        m.to(BF)
    params = train.set_trainable(p_unet, p_ref, p_proj, p_ad)
    opt = train.FlatAdamW(params, lr=1e-3, step_fn=emulated_ops.adamw_step)
    sched = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False)
    b = batch()
    train.train_step(sd, sched, optimizer=opt, **b)
    path = train.save_checkpoint(str(tmp_path), "ckpt-1", sd, opt, epoch=2, last_global_step=17, note="x")
    st = torch.load(path, map_location="cpu", weights_only=False)
    assert st["epoch"] == 2 and st["last_global_step"] == 17 and st["note"] == "x"
    keys = list(st["module"].keys())
    assert any(k.startswith("unet.") for k in keys) and any(k.startswith("ref_unet.") for k in keys)
    assert any(k.startswith("proj.") for k in keys) and any(k.startswith("adapter_modules.") and "to_k_ref" in k for k in keys)
    snap_p = opt.param.clone()
    snap_m = opt.m.clone()
    l2 = float(train.train_step(sd, sched, optimizer=opt, **b))  # moves on ...
    assert not torch.equal(opt.param, snap_p)
    epoch, step = train.load_checkpoint(str(tmp_path), sd, opt)  # ... and comes back (tag from `latest`)
    assert (epoch, step) == (2, 17) and opt.t == 1
    assert torch.equal(opt.param, snap_p) and torch.equal(opt.m, snap_m)
    assert p_ref.conv_in.weight.data_ptr() >= opt.param.data_ptr()  # still views of the flat buffer
    assert abs(float(train.train_step(sd, sched, optimizer=opt, **b)) - l2) < 1e-6  # the resumed step repeats the step it replaced
    # the inference-side routing of the reference reads the same file: `ref_unet.` keys load into a fresh garment UNet
    from imagdressing_b200 import modeling
    from test_train_cpu import CFG

    fresh = modeling.UNet2DConditionModel(**CFG)
    fresh.load_state_dict({k[len("ref_unet."):]: v.float() for k, v in st["module"].items() if k.startswith("ref_unet.")})


@pytest.mark.parametrize("name", ["linear", "cosine", "cosine_with_restarts", "polynomial", "constant", "constant_with_warmup"])
def test_lr_schedules_match_the_published_ones(name):
    """train.py:433-439 get_scheduler(args.lr_scheduler, optimizer, num_warmup_steps, num_training_steps) — diffusers 0.24.0
    is not in this image; its schedules are transformers' (same formulas, same defaults), which IS here and is the oracle."""
    from transformers.optimization import get_scheduler

    from imagdressing_b200 import train

    base, W, T = 1e-4, 3, 11
    lin = torch.nn.Linear(4, 4).to(BF)
    opt = train.FlatAdamW(lin.parameters(), lr=base, step_fn=emulated_ops.adamw_step)
    mine = train.LRScheduler(name, opt, num_warmup_steps=W, num_training_steps=T)
    ref_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=base)
    ref = get_scheduler(name, ref_opt, num_warmup_steps=W, num_training_steps=T)
    for step in range(T + 4):
        want = ref.get_last_lr()[0]
        assert mine.get_lr()[0] == pytest.approx(want, rel=1e-12, abs=1e-18), (name, step)
        assert opt.lr == mine.get_lr()[0] and float(opt.hyper[0]) == pytest.approx(want, rel=1e-6, abs=1e-12)  # device scalar
        ref_opt.step()
        ref.step()
        mine.step()
    resumed = train.LRScheduler(name, opt, num_warmup_steps=W, num_training_steps=T)
    resumed.load_state_dict(mine.state_dict())
    assert resumed.get_lr() == mine.get_lr() and resumed.last_epoch == T + 4
    with pytest.raises(ValueError):
        train.lr_multiplier("exponential", 0)


def test_gradient_accumulation_equals_the_whole_batch_step(monkeypatch):
    """train.py:606 `(step + 1) % gradient_accumulation_steps == 0`: k micro-batches, one update with the mean gradient."""
    emulated_ops.install(monkeypatch)
    from imagdressing_b200 import train

    def make(accum):
        torch.manual_seed(3)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4)).to(BF)
        return net, train.FlatAdamW(net.parameters(), lr=1e-2, weight_decay=0.01, bucket_bytes=32,
                                    step_fn=emulated_ops.adamw_step, accumulation_steps=accum)

    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1)).to(BF)
    whole_net, whole = make(1)
    acc_net, acc = make(4)
    assert float(acc.hyper[3]) == 0.25
    for _ in range(2):
        whole.zero_grad()
        whole_net(x).float().square().mean().backward()
        assert whole.step() is True
        before = acc.param.clone()
        for k in range(4):
            acc.zero_grad()
            acc_net(x[2 * k:2 * k + 2]).float().square().mean().backward()
            updated = acc.step()
            assert updated is (k == 3)
            if k < 3:
                assert torch.equal(acc.param, before)  # nothing moves inside the window
        assert not torch.equal(acc.param, before)
        g_acc, g_whole = acc.grad.float() / 4, whole.grad.float()  # the window's summed gradient, scaled as the kernel does
        assert float((g_acc - g_whole).norm() / g_whole.norm()) < 2e-2
    assert acc.t == 2 and whole.t == 2
    assert float((acc.param.float() - whole.param.float()).abs().max()) < 2e-2  # bf16 gradient sums vs one bf16 gradient
    with pytest.raises(ValueError):
        train.GraphedTrainStep(None, None, acc, {})
