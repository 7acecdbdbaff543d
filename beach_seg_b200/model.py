"""Mirror of the reference's `src/model.py`: `SegGptLoss` and `PromptModel` with the same method names, argument
meaning and attributes, running on the kernels of libbseg.so.  Lightning is not a dependency here: `PromptModel` is a
plain `torch.nn.Module` exposing the hooks the reference's scripts call (`post_init`, `create_trainable_params`,
`forward`, `prepare_prompt`, `create_palette`, `process_pred_masks`, `training_step`, `validation_step`,
`configure_optimizers`).  `training_step` returns a loss whose graph reaches the prompt parameters through the
CUDA backward (`bseg_backward_to_prompt`); `sync_prompt_grads` is the explicit data-parallel exchange that stands in
for Lightning's implicit DDP; `fit` is the 30-line loop that stands in for `Trainer.fit`
(src/train.py:97-115)."""
from __future__ import annotations

from typing import Any, Optional

import torch

from . import ops
from .augment import InferenceAug, TrainAug  # noqa: F401  (re-exported: the datamodule pipelines, src/data.py:195-234)
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


class PromptGradExchange:
    """Data-parallel exchange of the prompt gradients (Lightning's implicit DDP in the reference, src/train.py:96-107)
    without per-step packing: ONE persistent dense fp32 buffer [N_prompts, 3*H*W] whose rows ARE the `.grad` tensors of
    the prompt parameters (autograd accumulates into them in place), ONE average all-reduce over it per step (NCCL over
    NVLink on the GPU box, gloo in the CPU tests), and no device->host synchronisation on the gradient path: which
    prompts any rank drew this step is known from an all-gather of the (CPU-drawn) prompt indices that is issued on a
    side stream at the START of the step and has long finished when the backward is done.  Prompts no rank selected
    keep `grad = None`, so AdamW skips them exactly as it does at world size 1 -- with Lightning's default DDP the
    reference would instead fail on the unused parameters."""

    def __init__(self, params, group=None):
        import torch.distributed as dist

        self.params = list(params)
        self.group = group
        self.world = dist.get_world_size(group)
        p0 = self.params[0]
        self.per = p0.numel()
        self.buf = torch.zeros((len(self.params), self.per), dtype=torch.float32, device=p0.device)
        self.views = [self.buf[i].view(p.shape) for i, p in enumerate(self.params)]
        self.side = torch.cuda.Stream(device=p0.device) if p0.is_cuda else None
        self.event = torch.cuda.Event() if p0.is_cuda else None
        self._gathered = None
        self._host = None

    def begin_step(self, prompt_idx: torch.Tensor) -> None:
        """Call before the forward: zero the buffer, hang its rows on the parameters, start the index all-gather."""
        import torch.distributed as dist

        self.buf.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v
        idx = prompt_idx.to(torch.int64).flatten()
        n = idx.numel()
        if self.side is None:  # CPU (gloo) path of the tests
            out = torch.empty(self.world * n, dtype=torch.int64)
            dist.all_gather_into_tensor(out, idx.contiguous(), group=self.group)
            self._host = out
            return
        dev = self.buf.device
        if self._gathered is None or self._gathered.numel() != self.world * n:
            self._gathered = torch.empty(self.world * n, dtype=torch.int64, device=dev)
            self._host = torch.empty(self.world * n, dtype=torch.int64).pin_memory()
        pinned = idx.pin_memory()
        self.side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.side):
            idx_dev = pinned.to(dev, non_blocking=True)
            dist.all_gather_into_tensor(self._gathered, idx_dev, group=self.group)
            self._host.copy_(self._gathered, non_blocking=True)
            self.event.record(self.side)
        self._keep = (pinned, idx_dev)

    def finish_step(self) -> None:
        """Call after backward(): one average all-reduce over the buffer; untouched prompts get grad = None."""
        import torch.distributed as dist

        dist.all_reduce(self.buf, op=dist.ReduceOp.AVG if self.buf.is_cuda else dist.ReduceOp.SUM, group=self.group)
        if not self.buf.is_cuda:  # gloo has no AVG
            self.buf.div_(self.world)
        if self.event is not None:
            self.event.synchronize()  # recorded before the forward was even enqueued: no stall
        used = set(self._host.tolist())
        for i, p in enumerate(self.params):
            p.grad = self.views[i] if i in used else None


def allreduce_prompt_grads(params, group=None) -> None:
    """Stateless form of the exchange for callers that did not go through `PromptGradExchange.begin_step` (one
    collective over a dense [N_prompts, 3*H*W + 1] fp32 buffer; the last column counts the ranks that touched a prompt).
    Costs a pack / unpack and one host sync per step; `PromptModel.sync_prompt_grads` uses the persistent exchange."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or not params:
        return
    per = params[0].numel()
    buf = torch.zeros((len(params), per + 1), dtype=torch.float32, device=params[0].device)
    for i, p in enumerate(params):
        if p.grad is not None:
            buf[i, :per] = p.grad.reshape(-1)
            buf[i, per] = 1.0
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    world = dist.get_world_size(group)
    used = (buf[:, per] > 0).tolist()
    for i, p in enumerate(params):
        p.grad = (buf[i, :per] / world).reshape(p.shape).clone() if used[i] else None


class PromptModel(torch.nn.Module):
    def __init__(self, conf: BeachSegConfig, device: str | torch.device = "cuda:0", model=None):
        super().__init__()
        self.conf = conf
        self.num_classes = len(conf.classes)
        self.nodata_idx = 0
        # `model`: an already packed backbone to share (the reference always loads its own, src/model.py:77)
        self.model = model if model is not None else load_model(conf.checkpoint, device=device)
        self._device = torch.device(device)
        self.g = torch.Generator()  # the reference seeds a generator on model.device (cpu there)
        self.g.manual_seed(conf.seed)
        self.loss_fn = SegGptLoss(conf.loss_beta)
        self.aug: Any = InferenceAug()
        self.train_aug: Any = InferenceAug()
        self.prompt_batch: dict = {}
        self.prompt_params_list = torch.nn.ParameterList()
        self.dp_group = None  # torch.distributed group of the data-parallel ranks (None = the default group)

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
        """src/model.py:115-130: default-collate the prompt items (tensors stacked, strings listed), then replace
        "image" by the list of trainable Parameters."""
        from torch.utils.data import default_collate

        prompt_imgs = [{k: (torch.as_tensor(v) if not isinstance(v, (str, int, float)) else v) for k, v in p.items()}
                       for p in datamodule.prompt_imgs]
        self.prompt_batch = default_collate(prompt_imgs)
        params = [torch.nn.Parameter(self.prompt_batch["image"][i].to(self.device, torch.float32).clone(),
                                     requires_grad=True) for i in range(len(prompt_imgs))]
        self.prompt_params_list = torch.nn.ParameterList(params)
        self.prompt_batch["image"] = params

    # ---- src/model.py:132-147 ----
    @torch.no_grad()
    def forward(self, batch_dict: dict) -> torch.Tensor:
        B = batch_dict["image"].shape[0]
        batch_palette, batch_palette_norm = self.create_palette(B, train=True)
        prompt_batch, prompt_masks = self.prepare_prompt(batch_dict["crop_idx"], batch_palette, train=False)
        # opt-in: skip the decoder for the prompt half, whose pred_masks this method never reads (class maps identical)
        fast = {"query_half_only": True} if getattr(self, "query_half_only", False) else {}
        out = self.model(pixel_values=batch_dict["image"].to(self.device), prompt_pixel_values=prompt_batch["image"],
                         prompt_masks=prompt_masks, embedding_type="instance", **fast)
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
        # training: the stack keeps the graph to the selected prompt parameters (src/model.py:194)
        prompt_batch["image"] = torch.stack([p.to(self.device) if train else p.detach().to(self.device)
                                             for p in prompt_batch["image"]], dim=0)
        prompt_batch["mask"] = torch.stack([torch.as_tensor(m) for m in prompt_batch["mask"]], dim=0).to(self.device)
        prompt_batch = (self.train_aug if train else self.aug)(prompt_batch)
        prompt_color_mask_norm = ops.colorize_norm(prompt_batch["mask"], batch_palette.to(self.device))
        return prompt_batch, prompt_color_mask_norm

    # ---- src/model.py:215-231 ----
    def create_palette(self, batch_size: int, train: bool):
        return _create_palette(self.num_classes, batch_size, train, self.device)

    # ---- src/model.py:233-269 ----
    def training_step(self, batch: dict, batch_idx: int = 0) -> torch.Tensor:
        B = batch["mask"].shape[0]
        batch_palette, batch_palette_norm = self.create_palette(B, train=True)
        color_mask_norm = ops.colorize_norm(batch["mask"].to(self.device), batch_palette)
        prompt_idx = torch.randint(0, len(self.prompt_params_list), (B,), generator=self.g)
        ex = self._exchange()
        if ex is not None:
            ex.begin_step(prompt_idx)
        prompt_batch, prompt_masks = self.prepare_prompt(prompt_idx, batch_palette, train=True)
        out = self.model(pixel_values=batch["image"].to(self.device), labels=color_mask_norm,
                         prompt_pixel_values=prompt_batch["image"], prompt_masks=prompt_masks,
                         embedding_type="instance")
        self.last_pred = self.process_pred_masks(out.pred_masks.detach(), batch_palette_norm)  # the F1 metric's input
        self.last_prompt_idx = prompt_idx
        return self.loss_fn(out.pred_masks, color_mask_norm, (batch["mask"] != 0).to(self.device))

    # ---- src/model.py:271-308 ----
    @torch.no_grad()
    def validation_step(self, batch: dict, batch_idx: int = 0, dataloader_idx: int = 0) -> torch.Tensor:
        B = batch["mask"].shape[0]
        batch_palette, batch_palette_norm = self.create_palette(B, train=True)
        color_mask_norm = ops.colorize_norm(batch["mask"].to(self.device), batch_palette)
        prompt_batch, prompt_masks = self.prepare_prompt(batch["crop_idx"], batch_palette, train=False)
        out = self.model(pixel_values=batch["image"].to(self.device), labels=color_mask_norm,
                         prompt_pixel_values=prompt_batch["image"], prompt_masks=prompt_masks,
                         embedding_type="instance")
        self.last_pred = self.process_pred_masks(out.pred_masks, batch_palette_norm)
        return self.loss_fn(out.pred_masks, color_mask_norm, (batch["mask"] != 0).to(self.device))

    # ---- data-parallel exchange (Lightning's implicit DDP in the reference, src/train.py:96-107) ----
    def _exchange(self) -> Optional[PromptGradExchange]:
        """The persistent gradient exchange when a process group with more than one rank is up (else None)."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.dp_group) == 1 \
                or len(self.prompt_params_list) == 0:
            return None
        ex = getattr(self, "_grad_exchange", None)
        params = list(self.prompt_params_list)
        if ex is None or len(ex.params) != len(params) or any(a is not b for a, b in zip(ex.params, params)):
            ex = PromptGradExchange(params, self.dp_group)
            self._grad_exchange = ex
        return ex

    def sync_prompt_grads(self, group=None) -> None:
        """Mean of the prompt gradients over the data-parallel ranks; call between backward() and the optimiser step."""
        ex = self._exchange() if group is None or group is self.dp_group else None
        if ex is not None and ex._host is not None:
            ex.finish_step()
        else:
            allreduce_prompt_grads(list(self.prompt_params_list), group)

    # ---- src/model.py:385-428 (AdamW, optional linear warm-up, cosine per epoch; plain torch, negligible cost) ----
    def configure_optimizers(self):
        from torch.optim.lr_scheduler import CosineAnnealingLR, LambdaLR, SequentialLR

        global_bs = self.conf.batch_size * max(self.conf.world_size, 1) * self.conf.grad_accum_steps
        ratio = (global_bs / self.conf.base_lr_batch_size) ** 0.5
        lr, init_lr, min_lr = self.conf.lr * ratio, self.conf.init_lr * ratio, self.conf.min_lr * ratio
        warmup = self.conf.warmup_epochs
        if self.conf.optimizer != "adamw":
            raise RuntimeError(f"Unexpected optimizer {self.conf.optimizer}")
        opt = torch.optim.AdamW(self.parameters(), lr=lr)
        scheds, milestones = [], []
        if warmup:
            scheds.append(LambdaLR(opt, [lambda epoch: ((lr - init_lr) * (epoch / warmup) + init_lr) / lr]))
            milestones.append(warmup)
        if self.conf.scheduler != "cosine":
            raise RuntimeError(f"Unexpected scheduler {self.conf.scheduler}")
        scheds.append(CosineAnnealingLR(opt, self.conf.epochs, min_lr))
        return {"optimizer": opt, "lr_scheduler": {"scheduler": SequentialLR(opt, scheds, milestones),
                                                   "interval": "epoch", "frequency": 1}}

    # ---- stand-in for Trainer.fit (src/train.py:97-115): step, backward, exchange, AdamW, cosine per epoch ----
    def fit(self, batches, epochs: Optional[int] = None, on_step=None):
        """Default length = the reference's `Trainer(max_epochs=conf.epochs * len(prompt_batch))` (src/train.py:98;
        `prompt_batch` is the saved dict, so `len()` is its number of KEYS), while the cosine schedule keeps
        `T_max = conf.epochs` exactly as src/model.py:417 has it."""
        cfg = self.configure_optimizers()
        opt, sched = cfg["optimizer"], cfg["lr_scheduler"]["scheduler"]
        losses = []
        if epochs is None:
            epochs = self.conf.epochs * max(len(self.prompt_batch), 1)
        for _ in range(epochs):
            for i, batch in enumerate(batches):
                loss = self.training_step(batch, i)
                loss.backward()
                self.sync_prompt_grads()
                opt.step()
                opt.zero_grad(set_to_none=True)
                losses.append(loss.detach())
                if on_step is not None:
                    on_step(i, loss)
            sched.step()
        return torch.stack(losses).cpu() if losses else torch.empty(0)
