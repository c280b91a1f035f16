"""Image-file batch sources (list files and R-CNN window files) decoded with OpenCV.

reference: src/caffe/util/io.cpp:83-130 (ReadImageToDatum: BGR, optional resize),
src/caffe/layers/image_data_layer.cpp:24-102, src/caffe/layers/window_data_layer.cpp.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np
import torch


def read_image(path: str, new_h: int = 0, new_w: int = 0, color: bool = True) -> np.ndarray:
    """Decode to (C,H,W) uint8, BGR channel order like the reference (OpenCV)."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_COLOR if color else cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise IOError(f"Could not open or find file {path}")
    if new_h > 0 and new_w > 0:
        img = cv2.resize(img, (new_w, new_h))
    if img.ndim == 2:
        img = img[:, :, None]
    return np.ascontiguousarray(img.transpose(2, 0, 1))


def read_list_file(path: str) -> List[Tuple[str, int]]:
    root = os.path.dirname(os.path.abspath(path))
    out = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            name, label = line.rsplit(None, 1)
            if not os.path.isabs(name) and not os.path.exists(name):
                name = os.path.join(root, name)
            out.append((name, int(label)))
    return out


class ImageListSource:
    def __init__(self, list_file, batch, new_h=0, new_w=0, shuffle=False, rand_skip=0,
                 offset=0, stride=1, seed=None):
        self.lines = read_list_file(list_file)
        self.rng = np.random.RandomState(seed if seed is not None and seed >= 0 else None)
        if shuffle:
            self.rng.shuffle(self.lines)
        self.shuffle = shuffle
        self.batch, self.new_h, self.new_w = batch, new_h, new_w
        self.stride = max(1, stride)
        self.pos = offset % len(self.lines)
        if rand_skip:
            self.pos = (self.pos + int(self.rng.randint(0, rand_skip)) * self.stride) % len(self.lines)

    def next_batch(self):
        imgs, labels = [], []
        for _ in range(self.batch):
            name, label = self.lines[self.pos]
            imgs.append(read_image(name, self.new_h, self.new_w))
            labels.append(label)
            self.pos += self.stride
            if self.pos >= len(self.lines):
                self.pos %= len(self.lines)
                if self.shuffle:
                    self.rng.shuffle(self.lines)
        return torch.from_numpy(np.stack(imgs)), torch.tensor(labels, dtype=torch.float32)


class WindowSource:
    """Window file format:
        # image_index
        img_path
        channels height width
        num_windows
        class_index overlap x1 y1 x2 y2   (× num_windows)
    Each batch draws ``fg_fraction`` foreground windows (overlap ≥ fg_threshold) and the rest
    background (overlap < bg_threshold, label 0), each warped to crop×crop with optional
    context padding and random mirroring."""

    def __init__(self, window_file, batch, crop, fg_thr, bg_thr, fg_frac, context_pad, crop_mode,
                 mirror=False, seed=None):
        self.batch, self.crop = batch, crop
        self.fg_frac, self.context_pad, self.crop_mode, self.mirror = fg_frac, context_pad, crop_mode, mirror
        self.rng = np.random.RandomState(seed if seed is not None and seed >= 0 else None)
        self.images: List[str] = []
        self.fg, self.bg = [], []
        root = os.path.dirname(os.path.abspath(window_file))
        with open(window_file) as f:
            toks = f.read().split()
        i = 0
        while i < len(toks):
            assert toks[i] == "#", "bad window file"
            img_idx = int(toks[i + 1])
            path = toks[i + 2]
            if not os.path.isabs(path) and not os.path.exists(path):
                path = os.path.join(root, path)
            assert img_idx == len(self.images)
            self.images.append(path)
            nwin = int(toks[i + 6])
            i += 7
            for _ in range(nwin):
                cls, ov = int(toks[i]), float(toks[i + 1])
                box = tuple(int(float(t)) for t in toks[i + 2:i + 6])
                i += 6
                if ov >= fg_thr:
                    self.fg.append((img_idx, cls, box))
                elif ov < bg_thr:
                    self.bg.append((img_idx, 0, box))
        if not self.fg or not self.bg:
            raise ValueError("window file needs both foreground and background windows")

    def _crop(self, img, box, flip):
        import cv2
        c, h, w = img.shape
        x1, y1, x2, y2 = box
        size = self.crop
        pad_w = pad_h = 0
        out_w = out_h = size
        if self.context_pad > 0 or self.crop_mode == "square":
            scale = size / float(size - 2 * self.context_pad)
            half_h, half_w = (y2 - y1 + 1) / 2.0, (x2 - x1 + 1) / 2.0
            cx, cy = x1 + half_w, y1 + half_h
            if self.crop_mode == "square":
                half_h = half_w = max(half_h, half_w)
            x1, x2 = int(round(cx - half_w * scale)), int(round(cx + half_w * scale))
            y1, y2 = int(round(cy - half_h * scale)), int(round(cy + half_h * scale))
            uw, uh = x2 - x1 + 1, y2 - y1 + 1
            px1, py1 = max(0, -x1), max(0, -y1)
            px2, py2 = max(0, x2 - w + 1), max(0, y2 - h + 1)
            x1, x2, y1, y2 = x1 + px1, x2 - px2, y1 + py1, y2 - py2
            sx, sy = size / float(uw), size / float(uh)
            out_w, out_h = int(round((x2 - x1 + 1) * sx)), int(round((y2 - y1 + 1) * sy))
            pad_w = int(round((px2 if flip else px1) * sx))
            pad_h = int(round(py1 * sy))
            out_w, out_h = min(out_w, size - pad_w), min(out_h, size - pad_h)
        patch = img[:, max(0, y1):y2 + 1, max(0, x1):x2 + 1].transpose(1, 2, 0)
        patch = cv2.resize(np.ascontiguousarray(patch), (max(1, out_w), max(1, out_h)))
        if patch.ndim == 2:
            patch = patch[:, :, None]
        if flip:
            patch = patch[:, ::-1]
        out = np.zeros((c, size, size), dtype=np.uint8)
        out[:, pad_h:pad_h + patch.shape[0], pad_w:pad_w + patch.shape[1]] = patch.transpose(2, 0, 1)
        return out

    def next_batch(self):
        n_fg = int(self.batch * self.fg_frac)
        xs, ys = [], []
        for is_fg, count in ((False, self.batch - n_fg), (True, n_fg)):
            pool = self.fg if is_fg else self.bg
            for _ in range(count):
                img_idx, cls, box = pool[self.rng.randint(len(pool))]
                flip = bool(self.mirror and self.rng.randint(2))
                xs.append(self._crop(read_image(self.images[img_idx]), box, flip))
                ys.append(cls)
        return torch.from_numpy(np.stack(xs)), torch.tensor(ys, dtype=torch.float32)
