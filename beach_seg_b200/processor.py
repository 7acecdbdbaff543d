"""Drop-in for what `src/util/ml_util.py:16-17 load_processor()` returns on the src/predict_no_prompt.py path: the
subset of HF `SegGptImageProcessor` the reference calls (`preprocess`, `post_process_semantic_segmentation`,
`image_mean`, `image_std`; src/predict_no_prompt.py:240-246,283-301, src/data.py:191-193), running on the device
through libbseg.so.  Semantics follow transformers 5.5.0's torchvision-backend processor
(HF:image_processing_seggpt.py + image_processing_backends.py): bicubic-antialias resize on uint8, fused
rescale+normalise, palette mask colouring with NEAREST resize, and the palette-argmin post-processing."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops
from .ml_util import build_palette


class BatchFeature(dict):
    """Minimal stand-in for transformers.BatchFeature: a dict with attribute access and `.to()`."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, device):
        return BatchFeature({k: v.to(device) for k, v in self.items()})


class SegGptOutputLike:
    """Anything with a `.pred_masks` attribute works as `outputs` (the reference passes the model output object)."""

    def __init__(self, pred_masks):
        self.pred_masks = pred_masks


class SegGptImageProcessorB200:
    image_mean = list(ops.IMAGE_MEAN)
    image_std = list(ops.IMAGE_STD)
    size = {"height": 448, "width": 448}
    rescale_factor = 1 / 255

    def __init__(self, device: str | torch.device = "cuda:0"):
        self.device = torch.device(device)

    # ---- helpers -------------------------------------------------------------------------------------------
    def get_palette(self, num_labels: int):
        return build_palette(num_labels)

    def _stack_u8(self, images) -> tuple[torch.Tensor, bool]:
        """list of HWC uint8 numpy arrays / CHW uint8 tensors (or one batched array) -> uint8 device tensor."""
        if isinstance(images, (np.ndarray, torch.Tensor)) and images.ndim == 4:
            t = torch.as_tensor(images)
        else:
            if isinstance(images, (np.ndarray, torch.Tensor)):
                images = [images]
            t = torch.stack([torch.as_tensor(np.ascontiguousarray(im) if isinstance(im, np.ndarray) else im)
                             for im in images])
        if t.dtype != torch.uint8:
            raise ValueError("images must be uint8 (pixel values 0..255), like the crops the reference passes")
        channels_first = t.shape[1] == 3 and t.shape[-1] != 3
        return t.to(self.device), channels_first

    # ---- HF:image_processing_seggpt.py:134-215 -------------------------------------------------------------
    def preprocess(self, images=None, prompt_images=None, prompt_masks=None, num_labels: Optional[int] = None,
                   return_tensors: str = "pt", data_format: str = "channels_first", **kwargs) -> BatchFeature:
        if images is None and prompt_images is None and prompt_masks is None:
            raise ValueError("At least one of images, prompt_images, prompt_masks must be specified.")
        if return_tensors != "pt" or data_format not in ("channels_first", "ChannelDimension.FIRST"):
            raise NotImplementedError("only return_tensors='pt', data_format='channels_first' (what the reference "
                                      "passes, src/predict_no_prompt.py:244-245)")
        data = BatchFeature()
        if images is not None:
            t, cf = self._stack_u8(images)
            data["pixel_values"] = ops.preprocess_u8(t, channels_first=cf)
        if prompt_images is not None:
            t, cf = self._stack_u8(prompt_images)
            data["prompt_pixel_values"] = ops.preprocess_u8(t, channels_first=cf)
        if prompt_masks is not None:
            if isinstance(prompt_masks, (np.ndarray, torch.Tensor)) and prompt_masks.ndim == 2:
                prompt_masks = [prompt_masks]
            m = torch.stack([torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x)
                             for x in prompt_masks]).to(self.device)
            if m.ndim == 4:
                m = m.squeeze(1)
            if num_labels is None:
                raise NotImplementedError("segmentation-map prompt masks need num_labels (the reference always "
                                          "passes it, src/predict_no_prompt.py:243)")
            pal = torch.tensor(self.get_palette(num_labels), dtype=torch.uint8)
            data["prompt_masks"] = ops.colorize_resize_norm255(m, pal, 448)
        return data

    __call__ = preprocess

    # ---- HF:image_processing_seggpt.py:254-321 -------------------------------------------------------------
    def post_process_semantic_segmentation(self, outputs, target_sizes: Optional[Sequence] = None,
                                           num_labels: Optional[int] = None):
        masks = outputs.pred_masks
        if num_labels is None:
            raise NotImplementedError("num_labels=None (channel-mean decoding) is not used by the reference")
        pal = torch.tensor(self.get_palette(num_labels), dtype=torch.float32)
        B = masks.shape[0]
        if target_sizes is None:
            return list(ops.postprocess_semantic(masks.to(self.device), pal))
        out = []
        for i in range(B):
            h, w = target_sizes[i]
            if h != w:
                raise NotImplementedError("square target sizes only (the reference's crops are squares)")
            out.append(ops.postprocess_semantic(masks[i:i + 1].to(self.device), pal, out_size=int(h))[0])
        return out


def load_processor(checkpoint: str = "BAAI/seggpt-vit-large", device: str | torch.device = "cuda:0"):
    """src/util/ml_util.py:16-17.  The processor has no learned state: `checkpoint` only selects the preprocessing
    config, which is SegGPT's fixed 448x448 / ImageNet mean-std setup."""
    return SegGptImageProcessorB200(device)
