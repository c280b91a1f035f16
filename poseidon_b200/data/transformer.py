"""DataTransformer: (pixel − mean)·scale, random/centre crop, random mirror — batched, on
whatever device the batch lives on (the sm100 engine runs it as one CUDA kernel that also
converts uint8 NCHW → bf16 NHWC so raw bytes are all that cross PCIe).

reference: src/caffe/data_transformer.cpp:10-125 (per-datum CPU loop; mean file indexed at
the *uncropped* position; TRAIN = random crop + mirror, TEST = centre crop, no mirror).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from .. import proto as P


class DataTransformer:
    def __init__(self, tp, phase, device="cpu", seed: Optional[int] = None, model_dir=None,
                 allow_missing_mean: bool = False):
        self.scale = float(tp.scale)
        self.mirror = bool(tp.mirror)
        self.crop = int(tp.crop_size)
        self.phase = phase
        self.device = torch.device(device)
        self.mean = None          # (C,H,W) tensor
        self.mean_values = None   # (C,) tensor
        if tp.has("mean_file"):
            from ..utils.paths import resolve
            path = resolve(tp.mean_file, model_dir)
            if os.path.exists(path):
                blob = P.read_binary(path, P.BlobProto)
                self.mean = torch.from_numpy(P.blob_to_array(blob)[0].copy()).to(self.device)
            elif os.environ.get("POSEIDON_SYNTHETIC_DATA", "0") == "1" or allow_missing_mean:
                import logging
                logging.getLogger("poseidon_b200").warning(
                    "mean_file %s not found; using zero mean (synthetic data explicitly allowed)", path)
            else:
                # reference: data_transformer.cpp:19-27 CHECKs that the mean file can be read
                raise IOError(f"mean_file {path} not found (set POSEIDON_SYNTHETIC_DATA=1 / --synthetic_shape to run "
                              "without it on stand-in data)")
        if len(tp.mean_value):
            if tp.has("mean_file"):
                raise ValueError("Cannot specify mean_file and mean_value at the same time")
            self.mean_values = torch.tensor(list(tp.mean_value), dtype=torch.float32, device=self.device)
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(seed if seed is not None and seed >= 0 else
                             int.from_bytes(os.urandom(4), "little"))

    def out_hw(self, h, w):
        return (self.crop, self.crop) if self.crop else (h, w)

    def draw(self, n, h, w):
        """Per-sample crop offsets and mirror flags (host RNG, as the reference's own RNG)."""
        train = self.phase == P.TRAIN
        if self.crop:
            if train:
                h_off = torch.randint(0, h - self.crop + 1, (n,), generator=self.gen)
                w_off = torch.randint(0, w - self.crop + 1, (n,), generator=self.gen)
            else:
                h_off = torch.full((n,), (h - self.crop) // 2, dtype=torch.long)
                w_off = torch.full((n,), (w - self.crop) // 2, dtype=torch.long)
        else:
            h_off = torch.zeros(n, dtype=torch.long)
            w_off = torch.zeros(n, dtype=torch.long)
        if self.mirror and train:
            flip = torch.randint(0, 2, (n,), generator=self.gen).bool()
        else:
            flip = torch.zeros(n, dtype=torch.bool)
        return h_off, w_off, flip

    def __call__(self, x: torch.Tensor, out_dtype=torch.float32, draws=None) -> torch.Tensor:
        """x: (N,C,H,W) uint8 or float on any device -> (N,C,oh,ow) ``out_dtype``."""
        n, c, h, w = x.shape
        oh, ow = self.out_hw(h, w)
        if self.crop and (h < self.crop or w < self.crop):
            raise ValueError("crop_size larger than the input image")
        h_off, w_off, flip = draws if draws is not None else self.draw(n, h, w)
        dev = x.device
        h_off, w_off, flip = h_off.to(dev), w_off.to(dev), flip.to(dev)
        rows = h_off[:, None] + torch.arange(oh, device=dev)[None, :]          # (N, oh)
        cols = torch.arange(ow, device=dev)[None, :].expand(n, ow)
        cols = torch.where(flip[:, None], ow - 1 - cols, cols) + w_off[:, None]  # (N, ow)
        ni = torch.arange(n, device=dev)[:, None, None, None]
        ci = torch.arange(c, device=dev)[None, :, None, None]
        ri = rows[:, None, :, None]
        wi = cols[:, None, None, :]
        y = x[ni, ci, ri, wi].float()
        if self.mean is not None:
            m = self.mean.to(dev)
            if m.shape[0] != c or m.shape[1] != h or m.shape[2] != w:
                raise ValueError("mean file shape does not match the data")
            y = y - m[ci, ri, wi]
        elif self.mean_values is not None:
            mv = self.mean_values.to(dev)
            if mv.numel() == 1:
                mv = mv.expand(c)
            y = y - mv.view(1, c, 1, 1)
        if self.scale != 1.0:
            y = y * self.scale
        return y.to(out_dtype)
