"""Mirror of the reference's `src/model.py`: `SegGptLoss` and `PromptModel` with the same method names, argument
meaning and attributes, running on the kernels of libbseg.so.  Lightning is not a dependency here: `PromptModel` is a
plain `torch.nn.Module` exposing the hooks the reference's scripts call (`post_init`, `create_trainable_params`,
`forward`, `prepare_prompt`, `create_palette`, `process_pred_masks`, `training_step`, `configure_optimizers`).

Not built yet: the gradient of the backbone w.r.t. the prompt pixels, so `training_step`
computes the forward loss but the loss does not carry a graph to the prompt parameters (it raises if asked to)."""
from __future__ import annotations

from typing import Any, Optional

import torch

from . import ops
from .config import BeachSegConfig
from .ml_util import load_model
from .predict import create_palette as _create_palette


class _SmoothL1Fn(torch.autograd.Function):
    """loss(pred_masks) with d(loss)/d(pred_masks) from the fused forward+backward kernel."""

    @staticmethod
    def forward(ctx, pred_masks, labels, yesdata, beta, per_sample):
        loss, grad = ops.smooth_l1_loss(pred_masks.detach(), labels, yesdata, beta, per_sample=per_sample,
                                        want_grad=True)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None


class SegGptLoss(torch.nn.Module):
    """src/model.py:40-64.  `per_sample=False` (default) is the reference as written, including its BxB broadcast of
    the keep mask at batch > 1; `per_sample=True` is the B=1-equivalent form."""

    def __init__(self, beta: float, per_sample: bool = False):
        super().__init__()
        self.beta = beta
        self.per_sample = per_sample

    def forward(self, pred_masks: torch.Tensor, labels: torch.Tensor, yesdata: torch.Tensor) -> torch.Tensor:
        if pred_masks.requires_grad:
            return _SmoothL1Fn.apply(pred_masks, labels, yesdata, self.beta, self.per_sample)
        return ops.smooth_l1_loss(pred_masks, labels, yesdata, self.beta, per_sample=self.per_sample)


class InferenceAug:
    """Stand-in for the reference datamodule's `aug` pipeline (src/data.py:226-234): CenterCrop(inpt_size) is the
    identity on inpt_size inputs, Normalize(mean, std) is applied to the image entry; masks pass through."""

    def __call__(self, batch: dict) -> dict:
        out = dict(batch)
        mean = torch.tensor(ops.IMAGE_MEAN, dtype=torch.float32, device=batch["image"].device).view(1, 3, 1, 1)
        std = torch.tensor(ops.IMAGE_STD, dtype=torch.float32, device=batch["image"].device).view(1, 3, 1, 1)
        out["image"] = (batch["image"] - mean) / std
        return out

    def __len__(self):
        return 2


class PromptModel(torch.nn.Module):
    def __init__(self, conf: BeachSegConfig, device: str | torch.device = "cuda:0"):
        super().__init__()
        self.conf = conf
        self.num_classes = len(conf.classes)
        self.nodata_idx = 0
        self.model = load_model(conf.checkpoint, device=device)
        self._device = torch.device(device)
        self.g = torch.Generator()  # the reference seeds a generator on model.device (cpu there)
        self.g.manual_seed(conf.seed)
        self.loss_fn = SegGptLoss(conf.loss_beta)
        self.aug: Any = InferenceAug()
        self.train_aug: Any = InferenceAug()
        self.prompt_batch: dict = {}
        self.prompt_params_list = torch.nn.ParameterList()

    @property
    def device(self) -> torch.device:
        return self._device

    # ---- src/model.py:104-130 ----
    def post_init(self, datamodule: Any):
        self.train_aug = datamodule.train_aug
        self.aug = datamodule.aug

    def normalize(self, x: torch.Tensor) -> torch.Tensor:
        mean = torch.tensor(ops.IMAGE_MEAN, dtype=torch.float32, device=x.device).view(1, 3, 1, 1)
        std = torch.tensor(ops.IMAGE_STD, dtype=torch.float32, device=x.device).view(1, 3, 1, 1)
        return (x - mean) / std

    def denormalize(self, x: torch.Tensor) -> torch.Tensor:
        mean = torch.tensor(ops.IMAGE_MEAN, dtype=torch.float32, device=x.device).view(1, 3, 1, 1)
        std = torch.tensor(ops.IMAGE_STD, dtype=torch.float32, device=x.device).view(1, 3, 1, 1)
        return x * std + mean

    def create_trainable_params(self, datamodule: Any):
        prompt_imgs = datamodule.prompt_imgs  # list of dicts like BeachSegDataset.get_crop returns
        self.prompt_batch = {k: [torch.as_tensor(p[k]) for p in prompt_imgs] for k in prompt_imgs[0]
                             if k in ("image", "mask", "nodata", "crop_idx")}
        params = [torch.nn.Parameter(img.to(self.device, torch.float32), requires_grad=True)
                  for img in self.prompt_batch["image"]]
        self.prompt_params_list = torch.nn.ParameterList(params)
        self.prompt_batch["image"] = params

    # ---- src/model.py:132-147 ----
    @torch.no_grad()
    def forward(self, batch_dict: dict) -> torch.Tensor:
        B = batch_dict["image"].shape[0]
        batch_palette, batch_palette_norm = self.create_palette(B, train=True)
        prompt_batch, prompt_masks = self.prepare_prompt(batch_dict["crop_idx"], batch_palette, train=False)
        out = self.model(pixel_values=batch_dict["image"].to(self.device), prompt_pixel_values=prompt_batch["image"],
                         prompt_masks=prompt_masks, embedding_type="instance")
        return self.process_pred_masks(out.pred_masks, batch_palette_norm)

    # ---- src/model.py:155-175 ----
    def process_pred_masks(self, in_pred_masks: torch.Tensor, batch_palette_norm: torch.Tensor) -> torch.Tensor:
        return ops.decode_palette(in_pred_masks.to(self.device), batch_palette_norm.to(self.device))

    # ---- src/model.py:177-213 ----
    def prepare_prompt(self, batch_idxes, batch_palette: torch.Tensor, train: bool):
        if isinstance(batch_idxes, torch.Tensor):
            idx = batch_idxes.flatten().tolist()
        elif isinstance(batch_idxes, int):
            idx = [batch_idxes]
        else:
            idx = list(batch_idxes)
        prompt_batch = {k: [v[i] for i in idx] for k, v in self.prompt_batch.items()}
        prompt_batch["image"] = torch.stack([p.detach() for p in prompt_batch["image"]], dim=0).to(self.device)
        prompt_batch["mask"] = torch.stack([torch.as_tensor(m) for m in prompt_batch["mask"]], dim=0).to(self.device)
        prompt_batch = (self.train_aug if train else self.aug)(prompt_batch)
        prompt_color_mask_norm = ops.colorize_norm(prompt_batch["mask"], batch_palette.to(self.device))
        return prompt_batch, prompt_color_mask_norm

    # ---- src/model.py:215-231 ----
    def create_palette(self, batch_size: int, train: bool):
        return _create_palette(self.num_classes, batch_size, train, self.device)

    # ---- src/model.py:233-269 (forward + loss; the graph to the prompt parameters is not built yet) ----
    def training_step(self, batch: dict, batch_idx: int = 0) -> torch.Tensor:
        B = batch["mask"].shape[0]
        batch_palette, batch_palette_norm = self.create_palette(B, train=True)
        color_mask_norm = ops.colorize_norm(batch["mask"].to(self.device), batch_palette)
        prompt_idx = torch.randint(0, len(self.prompt_params_list), (B,), generator=self.g)
        prompt_batch, prompt_masks = self.prepare_prompt(prompt_idx, batch_palette, train=True)
        with torch.no_grad():
            out = self.model(pixel_values=batch["image"].to(self.device), labels=color_mask_norm,
                             prompt_pixel_values=prompt_batch["image"], prompt_masks=prompt_masks,
                             embedding_type="instance")
        return self.loss_fn(out.pred_masks, color_mask_norm, (batch["mask"] != 0).to(self.device))

    # ---- src/model.py:385-428 (AdamW + cosine per epoch; plain torch, negligible cost) ----
    def configure_optimizers(self, steps_per_epoch: Optional[int] = None):
        eff_bs = self.conf.batch_size * self.conf.grad_accum_steps * max(self.conf.world_size, 1)
        lr = self.conf.lr * (eff_bs / self.conf.base_lr_batch_size) ** 0.5
        opt = torch.optim.AdamW(self.parameters(), lr=lr)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=max(self.conf.epochs, 1),
                                                           eta_min=self.conf.min_lr)
        return {"optimizer": opt, "lr_scheduler": sched}
