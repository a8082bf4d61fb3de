"""Serving core behind the reference's ``POST /`` contract (api/api.py:52-82) — SURVEY.md §8(f) rank 4.

The reference endpoint reads ``{"image_path": ...}``, preprocesses the image (api.py:31-45: resize the short
side to 256, centre-crop 224, scale to [0, 1], normalise with the ImageNet mean / std), runs ``model(batch)``
and answers ``{"response": ..., "status": 200, "time": ...}``; its comment at api.py:66 says what a segmentation
model must do instead of the classification top-1: draw the mask from the output, save the mask image under a
path and return that path.  This module is that endpoint for the U-Nets of this package:

* :meth:`SegmentationService.build_input` — the same preprocessing (PIL + numpy; crop sizes are multiples of
  16 so the four pooling levels divide exactly);
* requests are grouped into **batch buckets** (1, 2, 4, ... ``max_batch``); each (bucket, H, W) shape is captured
  once as a CUDA graph (:class:`~.graph.GraphedInference`) and replayed with a single launch — short batches
  are zero-padded up to their bucket;
* the mask (sigmoid > 0.5 for one output channel, i.e. a threshold on the logits; arg-max over channels
  otherwise) is taken on the device, only the uint8 mask crosses PCIe, and it is written as a PNG;
* :func:`create_app` wraps the service in the reference's FastAPI route.  The handler is a plain function, so it
  is testable without a network.

There is no CPU fallback: the model must live on a CUDA device.
"""
from __future__ import annotations

import datetime
import os
import threading
import time
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from .graph import GraphedInference

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _buckets(max_batch: int) -> List[int]:
    out, b = [], 1
    while b < max_batch:
        out.append(b)
        b *= 2
    out.append(max_batch)
    return out


class SegmentationService:
    def __init__(self, model: nn.Module, out_dir: str, in_channels: int = 3, resize: int = 256, crop: int = 224,
                 max_batch: int = 8, threshold: float = 0.5, use_graphs: bool = True,
                 mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD):
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("SegmentationService serves a CUDA model only (no CPU fallback)")
        if crop % 16:
            raise ValueError("crop must be a multiple of 16 (four 2x pooling levels)")
        self.model = model.eval()
        self.out_dir = Path(out_dir)
        self.out_dir.mkdir(parents=True, exist_ok=True)
        self.in_channels, self.resize, self.crop = in_channels, resize, crop
        self.max_batch, self.threshold, self.use_graphs = max_batch, threshold, use_graphs
        self.mean = np.asarray(mean, dtype=np.float32)[:in_channels] if in_channels == 3 else \
            np.full((in_channels,), float(np.mean(mean)), dtype=np.float32)
        self.std = np.asarray(std, dtype=np.float32)[:in_channels] if in_channels == 3 else \
            np.full((in_channels,), float(np.mean(std)), dtype=np.float32)
        self.buckets = _buckets(max_batch)
        self._graphs: Dict[Tuple[int, int, int], GraphedInference] = {}
        self._lock = threading.Lock()           # one replay at a time: graphs own static buffers
        self.latencies_ms: List[float] = []

    # ------------------------------------------------------------------ api.py:31-45
    def build_input(self, image_path) -> torch.Tensor:
        """(C, crop, crop) float32 host tensor from an image file (or an HxW[xC] numpy array)."""
        if isinstance(image_path, np.ndarray):
            arr = image_path
        else:
            from PIL import Image
            img = Image.open(image_path)
            img = img.convert("RGB" if self.in_channels == 3 else "L")
            w, h = img.size
            s = self.resize / min(w, h)                     # transforms.Resize(256): short side -> 256
            img = img.resize((max(self.resize, int(round(w * s))), max(self.resize, int(round(h * s)))),
                             Image.BILINEAR)
            arr = np.asarray(img)
        if arr.ndim == 2:
            arr = arr[..., None]
        h, w = arr.shape[:2]
        if h < self.crop or w < self.crop:
            raise ValueError(f"image {h}x{w} is smaller than the {self.crop}x{self.crop} crop")
        top, left = int(round((h - self.crop) / 2.0)), int(round((w - self.crop) / 2.0))   # CenterCrop
        arr = arr[top:top + self.crop, left:left + self.crop, :self.in_channels]
        x = arr.astype(np.float32) / (255.0 if arr.dtype == np.uint8 else 1.0)            # ToTensor
        x = (x - self.mean) / self.std                                                      # Normalize
        return torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1)))

    # ------------------------------------------------------------------ batched forward
    def _bucket(self, n: int) -> int:
        for b in self.buckets:
            if n <= b:
                return b
        return self.max_batch

    def _forward(self, batch: torch.Tensor) -> torch.Tensor:
        """logits for a device batch whose size is a bucket; CUDA-graph replay when enabled."""
        if not self.use_graphs:
            with torch.no_grad():
                return self.model(batch)
        key = (batch.shape[0], batch.shape[-2], batch.shape[-1])
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = GraphedInference(self.model, batch)
        return g(batch)

    def predict_masks(self, inputs: Sequence[torch.Tensor]) -> List[np.ndarray]:
        """uint8 (H, W) masks (class indices; 0/1 for a single-channel head) for host tensors (C, H, W)."""
        out: List[np.ndarray] = []
        with self._lock:
            for start in range(0, len(inputs), self.max_batch):
                chunk = inputs[start:start + self.max_batch]
                b = self._bucket(len(chunk))
                host = torch.zeros((b,) + tuple(chunk[0].shape), dtype=torch.float32).pin_memory()
                for i, t in enumerate(chunk):
                    host[i].copy_(t)
                logits = self._forward(host.to(self.device, non_blocking=True))
                if isinstance(logits, (tuple, list)):
                    logits = logits[0]
                if logits.shape[1] == 1:                       # sigmoid(x) > thr  <=>  x > logit(thr)
                    thr = float(np.log(self.threshold / (1.0 - self.threshold)))
                    mask = (logits[:, 0] > thr).to(torch.uint8)
                else:
                    mask = logits.argmax(dim=1).to(torch.uint8)
                mask = mask[:len(chunk)].cpu().numpy()
                out.extend(mask[i] for i in range(len(chunk)))
        return out

    # ------------------------------------------------------------------ the POST / handler (api.py:52-82)
    def _save_mask(self, mask: np.ndarray, image_path) -> str:
        from PIL import Image
        stem = Path(str(image_path)).stem if not isinstance(image_path, np.ndarray) else f"array_{id(mask):x}"
        path = self.out_dir / f"{stem}_mask.png"
        scale = 255 if mask.max(initial=0) <= 1 else 1
        Image.fromarray((mask * scale).astype(np.uint8)).save(path)
        return str(path)

    def handle(self, json_post: dict) -> dict:
        """One request: ``{"image_path": ...}`` -> ``{"response": <mask path>, "status": 200, "time": ...}``."""
        return self.handle_many([json_post])[0]

    def handle_many(self, requests: Sequence[dict]) -> List[dict]:
        """Several queued requests served as one bucketed batch (the batching a server loop would do)."""
        t0 = time.perf_counter()
        answers: List[Optional[dict]] = [None] * len(requests)
        inputs, owners = [], []
        for i, req in enumerate(requests):
            image_path = req.get("image_path") if isinstance(req, dict) else None
            try:
                if image_path is None:
                    raise KeyError("image_path")
                inputs.append(self.build_input(image_path))
                owners.append((i, image_path))
            except Exception as e:   # noqa: BLE001 — a bad request must not take the batch down
                answers[i] = {"response": f"{type(e).__name__}: {e}", "status": 400, "time": _now()}
        if inputs:
            masks = self.predict_masks(inputs)
            for (i, image_path), mask in zip(owners, masks):
                answers[i] = {"response": self._save_mask(mask, image_path), "status": 200, "time": _now()}
        self.latencies_ms.append((time.perf_counter() - t0) * 1e3)
        return answers   # type: ignore[return-value]

    def latency_percentiles(self) -> Dict[str, float]:
        lat = sorted(self.latencies_ms)
        if not lat:
            return {}
        return {"n": len(lat), "p50_ms": lat[len(lat) // 2], "p99_ms": lat[min(len(lat) - 1, int(0.99 * len(lat)))]}


def _now() -> str:
    return datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S")


def create_app(service: SegmentationService):
    """The reference's FastAPI application (api/api.py:48-82) around a :class:`SegmentationService`."""
    from fastapi import FastAPI, Request

    app = FastAPI()

    @app.post("/")
    async def create_item(request: Request):
        json_post = await request.json()
        return service.handle(json_post)

    return app
