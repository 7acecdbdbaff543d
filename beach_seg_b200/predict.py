"""Tiled inference over a scene: the tensor part of src/predict.py:232-262 (prompted) batched over tiles, plus the
`Accumulator` vote stitcher (src/predict.py:55-159) kept on the device, and the owner-computes tile sharding used
when several GPUs work on one scene (inference tiles need no communication)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Optional, Sequence

import numpy as np
import torch

from . import ops
from .ml_util import build_palette, generate_random_rgb_palette
from .seggpt import SegGptB200


def create_palette(num_classes: int, batch_size: int, train: bool, device) -> tuple[torch.Tensor, torch.Tensor]:
    """PromptModel.create_palette (src/model.py:215-231): uint8 palette (B,C,3) and its normalised float copy.
    train=True draws the random palette from the global CPU generator exactly like the reference."""
    if train:
        batch_palette = generate_random_rgb_palette(num_classes, batch_size, "cpu")
    else:
        palette = torch.tensor(build_palette(num_classes - 1), dtype=torch.uint8)
        batch_palette = palette[None].repeat(batch_size, 1, 1)
    mean = torch.tensor(ops.IMAGE_MEAN, dtype=torch.float32)
    std = torch.tensor(ops.IMAGE_STD, dtype=torch.float32)
    palette_norm = (batch_palette.to(torch.float32) / 255 - mean) / std  # host logic: B*C*3 numbers
    return batch_palette.to(device), palette_norm.to(device)


def shard_tiles(n_tiles: int, rank: int, world_size: int) -> range:
    """Owner-computes sharding: rank r owns a contiguous block of the (row-major sorted) tile list, so each GPU
    stitches a spatial stripe of the scene.  No data-path collective is needed for inference."""
    per = (n_tiles + world_size - 1) // world_size
    return range(min(rank * per, n_tiles), min((rank + 1) * per, n_tiles))


def shard_rows(boxes, n_rows: int, rank: int, world_size: int) -> tuple[tuple[int, int], tuple[int, int]]:
    """What rank r has to hold of a scene of `n_rows` rows when the tile list `boxes` ((x0, y0, x1, y1) rows, sorted
    row-major) is sharded with `shard_tiles`: (rows its tiles read, rows it reduces for the scene-global statistics).
    The statistics rows are an even split of ALL rows (tiles need not cover the scene); both ranges are half-open and
    the first may be empty (0, 0) for a rank without tiles."""
    ids = shard_tiles(len(boxes), rank, world_size)
    per = (n_rows + world_size - 1) // world_size
    stat_rows = (min(rank * per, n_rows), min((rank + 1) * per, n_rows))
    if len(ids) == 0:
        return (0, 0), stat_rows
    ys = [(int(boxes[i][1]), int(boxes[i][3])) for i in ids]
    y0 = max(min(y for y, _ in ys), 0)
    y1 = min(max(y for _, y in ys), n_rows)
    return (y0, max(y1, y0)), stat_rows


def upload_scene_rows(scene_host: torch.Tensor, nodata_host: torch.Tensor, scene_dev: torch.Tensor,
                      nodata_dev: torch.Tensor, row_ranges) -> int:
    """Host -> device copy of the union of the half-open row ranges only (each band's rows are one contiguous chunk of
    the band-planar [4,Hs,Ws] scene).  Returns the bytes copied.  `scene_host` / `nodata_host` should be pinned."""
    lo = min(r[0] for r in row_ranges if r[1] > r[0])
    hi = max(r[1] for r in row_ranges if r[1] > r[0])
    for band in range(scene_host.shape[0]):   # one contiguous chunk per band: a strided copy_ would be staged on the host
        scene_dev[band, lo:hi].copy_(scene_host[band, lo:hi], non_blocking=True)
    nodata_dev[lo:hi].copy_(nodata_host[lo:hi], non_blocking=True)
    return int(scene_host[:, lo:hi].numel() * scene_host.element_size()
               + nodata_host[lo:hi].numel() * nodata_host.element_size())


class TilePredictor:
    """ingest -> SegGPT forward -> palette decode (+resize to crop size) for batches of tile boxes of one scene."""

    def __init__(self, model: SegGptB200, crop_size: int, num_classes: int = 4, random_palette: bool = True,
                 query_half_only: bool = False):
        """query_half_only: run the decoder for the query half only (`bseg_forward_query_half`); class maps are
        bit-identical because `process_pred_masks` reads pred_masks[:, :, 448:] only (src/model.py:158-160)."""
        self.model = model
        self.query_half_only = query_half_only
        self.crop_size = crop_size
        self.num_classes = num_classes
        self.random_palette = random_palette  # the reference's forward passes train=True (src/model.py:134)
        self.image_size = getattr(model, "image_size", 448)
        if self.image_size != 448 and self.image_size != crop_size:
            raise ValueError(f"native-resolution model (image_size={self.image_size}) needs crop_size == image_size")

    @torch.no_grad()
    def predict_tiles(self, scene_u16: torch.Tensor, nodata: torch.Tensor, stats: torch.Tensor, boxes: torch.Tensor,
                      prompt_images: torch.Tensor, prompt_masks: torch.Tensor,
                      palette: Optional[tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
        """boxes int32 [n,4]; prompt_images float32 [n,3,448,448] (normalised); prompt_masks uint8 [n,448,448] class
        ids.  Returns uint8 [n,crop,crop] class maps (nodata pixels are NOT zeroed, like src/predict.py)."""
        dev = self.model.device
        n = boxes.shape[0]
        if n == 0:  # e.g. a shard that owns no tile of a small scene
            return torch.empty((0, self.crop_size, self.crop_size), dtype=torch.uint8, device=dev)
        if palette is None:
            palette = create_palette(self.num_classes, n, self.random_palette, dev)
        pal_u8, pal_norm = palette
        tiles = ops.ingest_tiles(scene_u16, nodata, stats, boxes, self.crop_size, out_size=self.image_size)
        prompt_color = ops.colorize_norm(prompt_masks, pal_u8)
        out = self.model(pixel_values=tiles["image"], prompt_pixel_values=prompt_images, prompt_masks=prompt_color,
                         embedding_type="instance", query_half_only=self.query_half_only)
        return ops.decode_palette(out.pred_masks, pal_norm, out_size=self.crop_size, dtype=torch.uint8)


def predict_scene(predictor: "TilePredictor", scene_u16: torch.Tensor, nodata: torch.Tensor, boxes: torch.Tensor,
                  prompt_images: torch.Tensor, prompt_masks: torch.Tensor,
                  palette: Optional[tuple[torch.Tensor, torch.Tensor]] = None, rank: int = 0, world_size: int = 1,
                  tiles_per_launch: int = 64, group=None, stats: Optional[torch.Tensor] = None,
                  canvas: Optional[torch.Tensor] = None, reduce: bool = True):
    """The tile loop of src/predict.py:232-262 over one scene, batched and sharded: rank r predicts its contiguous block
    of the tile list (`shard_tiles`) in launches of `tiles_per_launch`, votes into its own packed-u32 canvas
    (`Accumulator`'s uint8 (H,W,4) counters), then ONE sum-reduce of the canvases to rank 0 (byte counters cannot carry
    below 256 votes per pixel) and `np.argmax` there (src/predict.py:100).  Tiles themselves need no communication.
    boxes int32 [n,4] (all tiles of the scene, every rank passes the same list); prompt_images [n,3,448,448];
    prompt_masks uint8 [n,448,448]; palette = (uint8 [n,C,3], float32 [n,C,3]) indexed by tile so that the result does
    not depend on the sharding.  Returns (class map uint8 [Hs,Ws] on rank 0 else None, this rank's canvas after the
    reduce)."""
    dev = predictor.model.device
    _, Hs, Ws = scene_u16.shape
    n = boxes.shape[0]
    if stats is None:
        stats = ops.scene_stats(scene_u16, nodata)  # scene-global, once per scene (src/util/geo_util.py:459-464)
    if canvas is None:
        canvas = torch.zeros((Hs, Ws), dtype=torch.int32, device=dev)
    else:
        canvas.zero_()
    if palette is None:
        palette = create_palette(predictor.num_classes, n, predictor.random_palette, dev)
    ids = shard_tiles(n, rank, world_size)
    for s in range(ids.start, ids.stop, tiles_per_launch):
        e = min(s + tiles_per_launch, ids.stop)
        cls = predictor.predict_tiles(scene_u16, nodata, stats, boxes[s:e], prompt_images[s:e], prompt_masks[s:e],
                                      (palette[0][s:e], palette[1][s:e]))
        ops.vote_accumulate(canvas, cls, boxes[s:e], overlapping=True)
    if world_size > 1 and reduce:
        import torch.distributed as dist

        dist.reduce(canvas, dst=0, op=dist.ReduceOp.SUM, group=group)
    pred = ops.vote_argmax(canvas) if rank == 0 else None
    return pred, canvas


class HostScenePipeline:
    """Host-buffer front end of `TilePredictor` for callers that hold the scene in (pinned) host memory, like the
    reference's numpy pipeline does: every `step()` copies the uint16 scene host->device, predicts its tiles, votes
    into the device canvas and copies the class maps back to a pinned host buffer.  The copies run on a side stream
    with two device / host buffers, so step i+1's upload and step i-1's download overlap step i's compute.  Uploads
    and downloads have a stream each: on one in-order copy stream the upload of step i+1 would queue behind the
    download of step i, which waits for step i's compute -- every upload would then be exposed."""

    def __init__(self, predictor: TilePredictor, scene_shape, n_tiles: int, crop_size: int):
        dev = predictor.model.device
        self.predictor, self.dev = predictor, dev
        self.h2d_stream = torch.cuda.Stream(device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)   # downloads (the stream a caller joins at the end)
        self.scene_dev = [torch.empty(scene_shape, dtype=torch.int16, device=dev) for _ in range(2)]
        self.cls_host = [torch.empty((n_tiles, crop_size, crop_size), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.h2d_done = [torch.cuda.Event() for _ in range(2)]
        self.buf_free = [torch.cuda.Event() for _ in range(2)]
        self.d2h_done = [torch.cuda.Event() for _ in range(2)]
        self.i = 0

    def step(self, scene_host: torch.Tensor, nodata, stats, boxes, prompt_images, prompt_cls, palette, canvas):
        """scene_host: pinned int16/uint16 [4,Hs,Ws].  Returns the pinned host tensor that will hold this step's class
        maps once `d2h_done[slot]` (also returned) has completed."""
        b = self.i % 2
        self.i += 1
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.h2d_stream):
            self.h2d_stream.wait_event(self.buf_free[b])       # the compute that last read this device buffer
            self.scene_dev[b].copy_(scene_host, non_blocking=True)
            self.h2d_done[b].record(self.h2d_stream)
        cur.wait_event(self.h2d_done[b])
        cls = self.predictor.predict_tiles(self.scene_dev[b], nodata, stats, boxes, prompt_images, prompt_cls, palette)
        ops.vote_accumulate(canvas, cls, boxes, overlapping=False)
        self.buf_free[b].record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.buf_free[b])
            cls.record_stream(self.copy_stream)
            self.cls_host[b].copy_(cls, non_blocking=True)
            self.d2h_done[b].record(self.copy_stream)
        return self.cls_host[b], self.d2h_done[b]

    def drain(self):
        self.h2d_stream.synchronize()
        self.copy_stream.synchronize()


class NoPromptPredictor:
    """The tensor part of src/predict_no_prompt.py:270-304, batched over tiles: HF-processor preprocessing of the uint8
    crops, `n_prompts` prompts per tile with `feature_ensemble=True` (ensemble grouped per tile), mean over the
    prompts, palette post-processing at crop size, nodata zeroing.  One launch handles `tiles x n_prompts` model
    samples."""

    def __init__(self, model: SegGptB200, processor, crop_size: int, num_classes: int = 4,
                 query_half_only: bool = False):
        self.model, self.processor, self.crop_size, self.num_classes = model, processor, crop_size, num_classes
        self.query_half_only = query_half_only  # post-processing reads pred_masks[:, :, 448:] only (HF:284-286)
        self._pal = torch.tensor(build_palette(num_classes - 1), dtype=torch.float32).to(model.device)

    @torch.no_grad()
    def predict_tiles(self, crops_u8: torch.Tensor, nodata: Optional[torch.Tensor], prompt_pixel_values: torch.Tensor,
                      prompt_masks: torch.Tensor) -> torch.Tensor:
        """crops_u8 uint8 [n,c,c,3]; nodata bool/uint8 [n,c,c] or None; prompt_pixel_values / prompt_masks float32
        [n*P,3,448,448] (the P prompts of tile i at rows i*P..i*P+P-1, already preprocessed).  Returns uint8 [n,c,c]."""
        n = crops_u8.shape[0]
        P = prompt_pixel_values.shape[0] // n
        px = ops.preprocess_u8(crops_u8.to(self.model.device))                      # :283-288
        px = px.repeat_interleave(P, dim=0)                                          # images=[crop_img] * len(prompts)
        out = self.model(pixel_values=px, prompt_pixel_values=prompt_pixel_values, prompt_masks=prompt_masks,
                         embedding_type="instance", feature_ensemble=True, ensemble_group=P,
                         query_half_only=self.query_half_only)                       # :289-295
        mean = ops.mean_over_prompts(out.pred_masks, P)                              # :298
        return ops.postprocess_semantic(mean, self._pal, out_size=self.crop_size, nodata=nodata,
                                        dtype=torch.uint8)                           # :299-303


def write_mask_tif(mask: np.ndarray, transform: Any, crs: Any, path: Path) -> bool:
    """src/util/img_util.py:67-95: single-band LZW GeoTIFF of the class map.  rasterio (GDAL) does the encoding; it is not
    part of this image, so without it (or without georeferencing) nothing is written and False is returned."""
    if transform is None or crs is None:
        return False
    try:
        import rasterio
    except ImportError:
        return False
    h, w = mask.shape
    with rasterio.open(path, "w", driver="GTiff", height=h, width=w, count=1, dtype=mask.dtype, transform=transform,
                       crs=crs, compress="lzw") as dst:
        dst.write(mask, 1)
    return True


class Accumulator:
    """Device-resident version of the reference's Accumulator (src/predict.py:55-159): same constructor and
    `update` / `save_current` / context-manager protocol.  `update` takes either the reference's one-hot uint8
    (crop,crop,C) array or a class-index map; votes are uint8 counters with the reference's wrap-around."""

    def __init__(self, out_shape: tuple[int, int], save_dir: Optional[Path], out_transform: Any = None,
                 crs: Any = None, classes: Sequence[str] = ("nodata", "sand", "water", "veg"),
                 device: str | torch.device = "cuda:0"):
        self.out_shape = tuple(out_shape)
        self.num_classes = len(classes)
        if self.num_classes > 4:
            raise ValueError("the packed vote counter holds at most 4 classes")
        self.out_transform, self.crs, self.classes = out_transform, crs, tuple(classes)
        self.device = torch.device(device)
        self.mask_dir = self.img_dir = self.tif_dir = None
        if save_dir is not None:  # the reference's output folders: images / masks / tif (src/predict.py:70-75)
            self.img_dir = Path(save_dir) / "images"
            self.mask_dir = Path(save_dir) / "masks"
            self.tif_dir = Path(save_dir) / "tif"
            for d in (self.img_dir, self.mask_dir, self.tif_dir):
                d.mkdir(exist_ok=True, parents=True)
        self.current_pred_counter: Optional[torch.Tensor] = None
        self.current_img: Optional[torch.Tensor] = None
        self.current_date = None

    def __enter__(self):
        assert self.current_pred_counter is None and self.current_date is None
        return self

    def __exit__(self, a, b, c):
        self.save_current()  # like the reference (src/predict.py:90-91): asserts if nothing was accumulated

    def initialize_current(self, date: str):
        self.current_date = date
        self.current_pred_counter = torch.zeros(self.out_shape, dtype=torch.int32, device=self.device)
        self.current_img = torch.zeros((*self.out_shape, 3), dtype=torch.uint8, device=self.device)

    def overlay(self) -> torch.Tensor:
        """overlay_prediction(current_img, argmax(counter), classes) (src/predict.py:100-102) as uint8 (H, W, 3) on the
        device."""
        assert self.current_img is not None
        return ops.overlay_prediction(self.current_img, self.prediction(), self.classes)

    def counter_u8(self) -> np.ndarray:
        """The reference's `current_pred_counter` view: uint8 (H, W, 4)."""
        assert self.current_pred_counter is not None
        return self.current_pred_counter.cpu().numpy().view(np.uint8).reshape(*self.out_shape, 4)

    def prediction(self) -> torch.Tensor:
        """np.argmax(counter, axis=2) (src/predict.py:100) as uint8 (H, W) on the device."""
        assert self.current_pred_counter is not None
        return ops.vote_argmax(self.current_pred_counter)

    def save_current(self):
        assert self.current_pred_counter is not None and self.current_date is not None
        pred_dev = self.prediction()
        pred = pred_dev.cpu().numpy()
        if self.mask_dir is not None:  # file encoders stay on the host; pixels come from the GPU
            import cv2
            from PIL import Image

            blended = ops.overlay_prediction(self.current_img, pred_dev, self.classes).cpu().numpy()
            Image.fromarray(blended).save(self.img_dir / f"{self.current_date}.png")
            cv2.imwrite(str(self.mask_dir / f"{self.current_date}.png"), pred)
            write_mask_tif(pred, self.out_transform, self.crs, self.tif_dir / f"{self.current_date}.tif")
        return pred

    def update(self, date: str, crop, one_hot_pred, img_crop=None, label_crop=None):
        if date != self.current_date:
            if self.current_pred_counter is not None:
                self.save_current()
            self.initialize_current(date)
        cls = one_hot_pred
        if isinstance(cls, np.ndarray):
            cls = torch.from_numpy(cls)
        if cls.ndim == 3 and cls.shape[-1] == self.num_classes and cls.shape[0] == cls.shape[1]:
            cls = cls.argmax(dim=-1)  # one-hot (crop,crop,C) -> index map (host/device view op, exact)
        cls = cls.to(device=self.device, dtype=torch.uint8)
        if cls.ndim == 2:
            cls = cls[None]
        boxes = torch.as_tensor(np.asarray(crop, dtype=np.int32).reshape(-1, 4), device=self.device)
        ops.vote_accumulate(self.current_pred_counter, cls.contiguous(), boxes)
        if img_crop is not None:
            img = torch.from_numpy(img_crop) if isinstance(img_crop, np.ndarray) else img_crop
            img = img.to(device=self.device, dtype=torch.uint8)
            ops.paste_tiles(self.current_img, img[None] if img.ndim == 3 else img, boxes)
