"""Training data for the IMAGDressing-v1 step: the record format, augmentation and batch layout of the
reference's IGPair dataset (/root/reference/IGPair.py:12-127), restated.

Record (one entry of the JSON list): {"image_file": person image path, "cloth_file": garment image path, "text": [captions]}.
Item:   person and garment images -> shorter side resized to 512 (bilinear) -> a random 640 x 512 crop -> [-1, 1] tensors
        (`vae_person`, `vae_clothes`); the garment image through the CLIP image processor (`clip_image`); one caption drawn at
        random, tokenised to the tokenizer's model_max_length (`text_input_ids`) beside the empty caption (`null_text_input_ids`);
        conditioning dropout (IGPair.py:60-68): with probability 0.05 the image embedding is dropped, 0.05 the caption, 0.05 both.
Batch (`collate_fn`, IGPair.py:104-127): stacked fp32 `vae_person` / `vae_clothes` [B, 3, 640, 512], `clip_image` [B, 3, 224, 224],
        `drop_image_embed` list, `text` list, `input_ids` / `null_input_ids` [B, T].
"""
from __future__ import annotations

import json
import random
from typing import Dict, List, Sequence, Union

import torch
from torch.utils.data import Dataset


class VDDataset(Dataset):
    """Same constructor and item keys as IGPair.VDDataset. `rng` (a random.Random) makes the augmentation reproducible in tests."""

    CROP = (640, 512)  # (height, width) of the training crop, IGPair.py:43-44

    def __init__(self, json_file: Union[str, Sequence[str]], tokenizer, size: int = 512, image_root_path: str = "",
                 clip_image_processor=None, rng: random.Random = None):
        if isinstance(json_file, str):
            files = [json_file]
        elif isinstance(json_file, (list, tuple)):
            files = list(json_file)
        else:
            raise ValueError("Input should be either a JSON file path (string) or a list")
        self.data: List[dict] = []
        for path in files:
            with open(path, "r", encoding="utf-8") as f:
                self.data.extend(json.load(f))
        self.tokenizer = tokenizer
        self.size = size
        self.image_root_path = image_root_path
        self.rng = rng or random
        if clip_image_processor is None:
            from transformers import CLIPImageProcessor

            clip_image_processor = CLIPImageProcessor()
        self.clip_image_processor = clip_image_processor

    def __len__(self) -> int:
        return len(self.data)

    def _vae_tensor(self, img) -> torch.Tensor:
        """Resize(512, bilinear) -> RandomCrop(640 x 512) -> ToTensor -> Normalize(0.5, 0.5): [3, 640, 512] in [-1, 1]."""
        from PIL import Image
        import numpy as np

        w, h = img.size
        s = 512.0 / min(w, h)
        nw, nh = max(512, round(w * s)), max(512, round(h * s))
        img = img.resize((nw, nh), Image.BILINEAR)
        ch, cw = self.CROP
        if nh < ch or nw < cw:  # torchvision's RandomCrop raises here too: the data set is portrait, >= 640 x 512 after resize
            raise ValueError(f"image {nw}x{nh} after resize is smaller than the {cw}x{ch} crop")
        top = self.rng.randint(0, nh - ch)
        left = self.rng.randint(0, nw - cw)
        arr = np.asarray(img.crop((left, top, left + cw, top + ch)), dtype=np.float32) / 255.0
        return torch.from_numpy(arr).permute(2, 0, 1).contiguous().sub_(0.5).div_(0.5)

    def _tokens(self, text: str) -> torch.Tensor:
        return self.tokenizer(text, max_length=self.tokenizer.model_max_length, padding="max_length", truncation=True,
                              return_tensors="pt").input_ids

    def __getitem__(self, idx: int) -> Dict[str, object]:
        from PIL import Image

        item = self.data[idx]
        person = Image.open(item["image_file"]).convert("RGB")
        clothes = Image.open(item["cloth_file"]).convert("RGB")
        text = self.rng.choice(item["text"])
        drop_image_embed = 0
        r = self.rng.random()
        if r < 0.05:
            drop_image_embed = 1
        elif r < 0.1:
            text = ""
        elif r < 0.15:
            text = ""
            drop_image_embed = 1
        return {
            "vae_person": self._vae_tensor(person),
            "vae_clothes": self._vae_tensor(clothes),
            "clip_image": self.clip_image_processor(images=clothes, return_tensors="pt").pixel_values,
            "drop_image_embed": drop_image_embed,
            "text": text,
            "text_input_ids": self._tokens(text),
            "null_text_input_ids": self._tokens(""),
        }


def collate_fn(data: List[dict]) -> Dict[str, object]:
    return {
        "vae_person": torch.stack([e["vae_person"] for e in data]).contiguous().float(),
        "vae_clothes": torch.stack([e["vae_clothes"] for e in data]).contiguous().float(),
        "clip_image": torch.cat([e["clip_image"] for e in data], dim=0),
        "drop_image_embed": [e["drop_image_embed"] for e in data],
        "text": [e["text"] for e in data],
        "input_ids": torch.cat([e["text_input_ids"] for e in data], dim=0),
        "null_input_ids": torch.cat([e["null_text_input_ids"] for e in data], dim=0),
    }
