"""Drop-in for the object `src/util/ml_util.py:7-13 load_model()` returns: the frozen HuggingFace
`SegGptForImageSegmentation` in eval mode, with the same call signature and output fields, running on the
hand-written sm_100a kernels of libbseg.so (HF:modeling_seggpt.py:839-959 for the semantics).

    model = SegGptB200.from_hf(hf_model)            # or load_model(checkpoint)
    out = model(pixel_values=..., prompt_pixel_values=..., prompt_masks=..., embedding_type="instance")
    out.pred_masks                                  # float32 [B, 3, 896, 448]; attribute is assignable
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib

IMG = 448
NUM_PATCHES = 1568


def _ops_loss_scratch() -> int:
    from .ops import LOSS_SCRATCH_FLOATS

    return LOSS_SCRATCH_FLOATS


class SegGptOutput:
    """Mutable stand-in for HF SegGptImageSegmentationOutput (src/predict_no_prompt.py:298 assigns .pred_masks)."""

    def __init__(self, loss=None, pred_masks=None):
        self.loss = loss
        self.pred_masks = pred_masks
        self.hidden_states = None
        self.attentions = None

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        return tuple(v for v in (self.loss, self.pred_masks) if v is not None)[k]


class _PromptGradFn(torch.autograd.Function):
    """pred_masks = SegGPT(pixel_values, prompt_pixel_values, prompt_masks) with a gradient path to
    prompt_pixel_values only: the reference trains nothing else (src/model.py:115-130, src/util/ml_util.py:9-10).
    Forward = bseg_forward_train (keeps activations in the module's training workspace), backward =
    bseg_backward_to_prompt.  One forward may be in flight per module (the workspace is reused)."""

    @staticmethod
    def forward(ctx, prompt_pixel_values, model, pixel_values, prompt_masks, embedding_type):
        B = pixel_values.shape[0]
        pred = torch.empty((B, 3, 2 * IMG, IMG), dtype=torch.float32, device=model.device)
        ws, base, nbytes = model._train_workspace(B)
        L = _lib.lib()
        fwd, name = ((L.bseg_forward_train_f32, "bseg_forward_train_f32") if model.precision == "fp32"
                     else (L.bseg_forward_train, "bseg_forward_train"))
        with torch.cuda.device(model.device):
            _lib.check(fwd(model._handle, _lib.ptr(pixel_values), _lib.ptr(prompt_pixel_values),
                           _lib.ptr(prompt_masks), B, 0 if embedding_type == "instance" else 1, C.c_void_p(base),
                           C.c_size_t(nbytes), _lib.ptr(pred), _lib.stream_ptr()), name)
        ctx.model, ctx.batch = model, B
        model._train_token += 1
        ctx.token = model._train_token
        return pred

    @staticmethod
    def backward(ctx, d_pred):
        model, B = ctx.model, ctx.batch
        if ctx.token != model._train_token:
            raise _lib.BsegError("backward() after another training forward of the same module: its saved "
                                 "activations were overwritten")
        d_pred = d_pred.to(torch.float32).contiguous()
        if model.check_grad_support and bool((d_pred[:, :, :IMG] != 0).any()):  # (a device->host sync)
            raise NotImplementedError("d(pred_masks) must be zero in the prompt half (image rows < 448), as it is for "
                                      "the reference's SegGptLoss (src/model.py:48-57)")
        d_prompt = torch.empty((B, 3, IMG, IMG), dtype=torch.float32, device=model.device)
        ws, base, nbytes = model._train_workspace(B)
        L = _lib.lib()
        bwd, name = ((L.bseg_backward_to_prompt_f32, "bseg_backward_to_prompt_f32") if model.precision == "fp32"
                     else (L.bseg_backward_to_prompt, "bseg_backward_to_prompt"))
        with torch.cuda.device(model.device):
            _lib.check(bwd(model._handle, _lib.ptr(d_pred), B, C.c_void_p(base), C.c_size_t(nbytes), _lib.ptr(d_prompt),
                           _lib.stream_ptr()), name)
        return d_prompt, None, None, None, None


class SegGptB200(torch.nn.Module):
    def __init__(self, state_dict: Dict[str, torch.Tensor], num_layers: int = 24, merge_index: int = 2,
                 intermediate=(5, 11, 17, 23), layer_norm_eps: float = 1e-6, beta: float = 0.01,
                 device: str | torch.device = "cuda:0", max_batch: int = 64, precision: str = "bf16",
                 graph_batch: int = 16, image_size: int = IMG):
        """precision: "bf16" = the tcgen05 path (bf16 operands, fp32 accumulation / residual stream / softmax; logits
        within 1e-2 of the fp32 reference); "fp32" = the accuracy mode (bseg_forward_f32: everything IEEE fp32 on the
        CUDA cores, within 1e-4, inference only, ~30x slower).
        image_size: 448 = the reference's path (crops resized to 448, `SegGptConfig()`); 512 / 1024 = native-resolution
        mode for 512- / 1024-px tiles (`SegGptConfig(image_size=(2 * tile, tile))`: 64 x 32 tokens, T = 2048, rel-pos tables
        of 127 / 63 rows; 128 x 64 tokens, T = 8192, 255 / 127 rows) -- bf16 inference only.
        graph_batch: inference calls with at most this many samples go through persistent staging buffers and a CUDA graph
        of the whole forward (bseg_set_graph_batch_limit): the reference's own call pattern is batch 1
        (src/predict.py:234), where the host work of ~190 launches is as long as the device work.  0 disables it."""
        super().__init__()
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.precision = precision
        if image_size not in (448, 512, 1024):
            raise ValueError(f"image_size must be 448, or 512 / 1024 (native-resolution tiles), got {image_size}")
        if image_size != IMG and precision != "bf16":
            raise NotImplementedError("the native-resolution mode runs on the bf16 path only")
        self.image_size = int(image_size)
        self.num_patches = 2 * (self.image_size // 16) ** 2
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise _lib.BsegError("SegGptB200 runs on a CUDA device only; there is no CPU path")
        self.num_layers, self.merge_index, self.intermediate = num_layers, merge_index, tuple(intermediate)
        self.beta = beta
        self.max_batch = max_batch
        self._handle = C.c_void_p()
        self._ws: Optional[torch.Tensor] = None
        self._train_ws: Optional[torch.Tensor] = None
        self._train_token = 0
        self.check_grad_support = True  # verify in backward() that d(pred_masks) is zero in the prompt half
        self._train_ready = False
        self._scratch = torch.zeros(_ops_loss_scratch(), dtype=torch.float32, device=self._device)
        self.graph_batch = int(graph_batch) if precision == "bf16" else 0
        self._stage: Dict[int, tuple] = {}  # batch -> (px, ppx, pm, pred, workspace) with fixed addresses
        L = _lib.lib()
        with torch.cuda.device(self._device):
            keep = []  # fp32 staging copies, freed after packing

            def dev(name):
                t = state_dict[name].detach().to(device=self._device, dtype=torch.float32).contiguous()
                keep.append(t)
                return C.c_void_p(t.data_ptr())

            layers = (_lib.LayerWeights * num_layers)()
            for i in range(num_layers):
                p = f"model.encoder.layers.{i}."
                lw = layers[i]
                lw.ln1_w, lw.ln1_b = dev(p + "layernorm_before.weight"), dev(p + "layernorm_before.bias")
                lw.qkv_w, lw.qkv_b = dev(p + "attention.qkv.weight"), dev(p + "attention.qkv.bias")
                lw.rel_pos_h, lw.rel_pos_w = dev(p + "attention.rel_pos_h"), dev(p + "attention.rel_pos_w")
                lw.proj_w, lw.proj_b = dev(p + "attention.proj.weight"), dev(p + "attention.proj.bias")
                lw.ln2_w, lw.ln2_b = dev(p + "layernorm_after.weight"), dev(p + "layernorm_after.bias")
                lw.lin1_w, lw.lin1_b = dev(p + "mlp.lin1.weight"), dev(p + "mlp.lin1.bias")
                lw.lin2_w, lw.lin2_b = dev(p + "mlp.lin2.weight"), dev(p + "mlp.lin2.bias")
            w = _lib.Weights()
            w.image_size = self.image_size
            w.num_layers, w.merge_index = num_layers, merge_index
            w.intermediate_indices = (C.c_int * 4)(*self.intermediate)
            w.layer_norm_eps = layer_norm_eps
            e = "model.embeddings."
            w.patch_w, w.patch_b = dev(e + "patch_embeddings.projection.weight"), dev(e + "patch_embeddings.projection.bias")
            w.mask_token = dev(e + "mask_token")
            w.segment_token_input, w.segment_token_prompt = dev(e + "segment_token_input"), dev(e + "segment_token_prompt")
            w.type_token_semantic, w.type_token_instance = dev(e + "type_token_semantic"), dev(e + "type_token_instance")
            w.position_embeddings = dev(e + "position_embeddings")
            w.layers = layers
            w.enc_ln_w, w.enc_ln_b = dev("model.encoder.layernorm.weight"), dev("model.encoder.layernorm.bias")
            w.dec_embed_w, w.dec_embed_b = dev("decoder.decoder_embed.weight"), dev("decoder.decoder_embed.bias")
            w.dec_conv_w, w.dec_conv_b = dev("decoder.decoder_pred.conv.weight"), dev("decoder.decoder_pred.conv.bias")
            w.dec_ln_w, w.dec_ln_b = dev("decoder.decoder_pred.layernorm.weight"), dev("decoder.decoder_pred.layernorm.bias")
            w.dec_head_w, w.dec_head_b = dev("decoder.decoder_pred.head.weight"), dev("decoder.decoder_pred.head.bias")
            _lib.check(L.bseg_create(C.byref(w), C.byref(self._handle), _lib.stream_ptr()), "bseg_create")
            if precision == "fp32":
                _lib.check(L.bseg_enable_fp32(self._handle, C.byref(w), _lib.stream_ptr()), "bseg_enable_fp32")
            if self.graph_batch > 0:
                _lib.check(L.bseg_set_graph_batch_limit(self._handle, self.graph_batch) < 0, "bseg_set_graph_batch_limit")
            torch.cuda.current_stream().synchronize()
            del keep

    # ------------------------------------------------------------------------------------------------
    @classmethod
    def from_hf(cls, hf_model, device="cuda:0", **kw) -> "SegGptB200":
        cfg = hf_model.config
        size = cfg.image_size
        height, width = (size, size) if isinstance(size, int) else (int(size[0]), int(size[1]))
        if height != 2 * width:
            raise ValueError(f"SegGptConfig.image_size={size}: the stacked image must be (2 * tile, tile)")
        kw.setdefault("image_size", width)
        return cls(hf_model.state_dict(), num_layers=cfg.num_hidden_layers, merge_index=cfg.merge_index,
                   intermediate=tuple(cfg.intermediate_hidden_state_indices), layer_norm_eps=cfg.layer_norm_eps,
                   beta=cfg.beta, device=device, **kw)

    @property
    def device(self) -> torch.device:
        return self._device

    def to(self, *args, **kwargs):  # the backbone lives in the library's arena; only cuda is meaningful
        return self

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                _lib.lib().bseg_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    def _workspace(self, batch: int) -> torch.Tensor:
        L = _lib.lib()
        need = int((L.bseg_workspace_bytes_f32 if self.precision == "fp32" else L.bseg_workspace_bytes)(
            self._handle, batch))
        if self._ws is None or self._ws.numel() < need + 256:
            self._ws = None
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self._device)
        return self._ws

    def _train_workspace(self, batch: int):
        """(tensor, 256B-aligned base address, usable bytes) of the training workspace for `batch` samples; the first
        call also packs the transposed weight copies the dgrad GEMMs need."""
        L = _lib.lib()
        if self.precision == "fp32":  # (the accuracy mode differentiates with the weights as stored: nothing to pack)
            need = int(L.bseg_train_workspace_bytes_f32(self._handle, batch))
        else:
            if not self._train_ready:
                with torch.cuda.device(self._device):
                    _lib.check(L.bseg_train_prepare(self._handle, _lib.stream_ptr()), "bseg_train_prepare")
                self._train_ready = True
            need = int(L.bseg_train_workspace_bytes(self._handle, batch))
        if self._train_ws is None or self._train_ws.numel() < need + 256:
            self._train_ws = None
            self._train_ws = torch.empty(need + 256, dtype=torch.uint8, device=self._device)
        base = (self._train_ws.data_ptr() + 255) // 256 * 256
        return self._train_ws, base, self._train_ws.numel() - (base - self._train_ws.data_ptr())

    # ------------------------------------------------------------------------------------------------
    def forward(self, pixel_values: torch.Tensor, prompt_pixel_values: torch.Tensor, prompt_masks: torch.Tensor,
                bool_masked_pos: Optional[torch.Tensor] = None, feature_ensemble: Optional[bool] = None,
                embedding_type: Optional[str] = None, labels: Optional[torch.Tensor] = None,
                output_attentions: Optional[bool] = None, output_hidden_states: Optional[bool] = None,
                return_dict: Optional[bool] = None, ensemble_group: Optional[int] = None,
                query_half_only: bool = False, **kwargs) -> SegGptOutput:
        """HF forward signature (HF:modeling_seggpt.py:839-959) plus two extensions: `ensemble_group` (several tiles of P
        prompts per launch) and `query_half_only` (skip the decoder for the prompt half: pred_masks[:, :, :448] is zero,
        the bottom half -- the only part the reference ever reads -- is bit-identical)."""
        S, NP = self.image_size, self.num_patches
        for name, t in (("pixel_values", pixel_values), ("prompt_pixel_values", prompt_pixel_values),
                        ("prompt_masks", prompt_masks)):
            if t.ndim != 4 or t.shape[1] != 3:
                raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in "
                                 "the configuration.")
            if t.shape[2] != S or t.shape[3] != S:
                # HF checks the stacked image (HF:modeling_seggpt.py:116-119)
                raise ValueError(f"Input image size ({2 * t.shape[2]}*{t.shape[3]}) doesn't match model ({2 * S}*{S}).")
        embedding_type = embedding_type if embedding_type is not None else "instance"
        if embedding_type not in ("instance", "semantic"):
            raise ValueError(f"Embedding type should be either 'semantic' or 'instance', but got {embedding_type}")
        mask_rows = 1  # HF's default bool_masked_pos has batch dimension 1 (HF:modeling_seggpt.py:910-917)
        if bool_masked_pos is not None:
            default = torch.cat([torch.zeros(NP // 2, dtype=torch.bool), torch.ones(NP - NP // 2, dtype=torch.bool)])
            rows = bool_masked_pos.reshape(-1, NP)
            if not torch.equal(rows.cpu().bool(), default.expand(rows.shape[0], -1)):
                raise NotImplementedError("only the default bool_masked_pos (bottom half masked) is supported; the "
                                          "reference never passes another one")
            mask_rows = rows.shape[0]
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention / hidden-state outputs are never requested on the reference path")
        want_grad = torch.is_grad_enabled() and prompt_pixel_values.requires_grad
        if torch.is_grad_enabled() and (pixel_values.requires_grad or prompt_masks.requires_grad):
            raise NotImplementedError("only prompt_pixel_values carries a gradient on the reference path "
                                      "(src/model.py:115-130); pixel_values / prompt_masks must not require grad")
        B = pixel_values.shape[0]
        if not (prompt_pixel_values.shape[0] == B and prompt_masks.shape[0] == B):
            raise ValueError("pixel_values, prompt_pixel_values and prompt_masks must share the batch dimension")
        P = 0
        if feature_ensemble:
            P = B if ensemble_group is None else int(ensemble_group)
            if P <= 0 or B % P != 0:
                raise ValueError(f"ensemble_group={P} does not divide the batch {B}")

        def prep(t):
            return t.detach().to(device=self._device, dtype=torch.float32).contiguous()

        if want_grad and S != IMG:
            raise NotImplementedError("the native-resolution mode is inference-only; the train step runs at 448")
        if want_grad:
            if feature_ensemble:
                raise NotImplementedError("feature_ensemble is an inference-only path in the reference "
                                          "(src/predict_no_prompt.py:289-295)")
            ppx_g = prompt_pixel_values.to(device=self._device, dtype=torch.float32).contiguous()
            pred = _PromptGradFn.apply(ppx_g, self, prep(pixel_values), prep(prompt_masks), embedding_type)
            return SegGptOutput(loss=self._hf_loss(pred.detach(), labels, B, mask_rows), pred_masks=pred)
        L = _lib.lib()
        if 0 < B <= min(self.graph_batch, self.max_batch) and self.precision == "bf16":
            return self._forward_staged(pixel_values, prompt_pixel_values, prompt_masks, B, embedding_type, P,
                                        query_half_only, labels, mask_rows)
        px, ppx, pm = prep(pixel_values), prep(prompt_pixel_values), prep(prompt_masks)
        pred = torch.empty((B, 3, 2 * S, S), dtype=torch.float32, device=self._device)
        step = self.max_batch if P == 0 else max(P, (self.max_batch // P) * P)
        if self.precision == "fp32":
            fwd, fwd_name = L.bseg_forward_f32, "bseg_forward_f32"
        elif query_half_only:
            fwd, fwd_name = L.bseg_forward_query_half, "bseg_forward_query_half"
        else:
            fwd, fwd_name = L.bseg_forward, "bseg_forward"
        with torch.cuda.device(self._device):
            for s in range(0, B, step):
                n = min(step, B - s)
                ws = self._workspace(n)
                base = (ws.data_ptr() + 255) // 256 * 256
                _lib.check(fwd(self._handle, _lib.ptr(px[s:s + n]), _lib.ptr(ppx[s:s + n]),
                               _lib.ptr(pm[s:s + n]), n, 0 if embedding_type == "instance" else 1, P,
                               C.c_void_p(base), C.c_size_t(ws.numel() - (base - ws.data_ptr())),
                               _lib.ptr(pred[s:s + n]), _lib.stream_ptr()), fwd_name)
        return SegGptOutput(loss=self._hf_loss(pred, labels, B, mask_rows), pred_masks=pred)

    def _forward_staged(self, pixel_values, prompt_pixel_values, prompt_masks, B, embedding_type, P, query_half_only,
                        labels, mask_rows) -> SegGptOutput:
        """Small-batch inference: inputs are copied into per-batch-size buffers with fixed addresses, so that the
        library can replay the whole forward as one CUDA graph (same kernels, same order: bit-identical results)."""
        st = self._stage.get(B)
        if st is None:
            L = _lib.lib()
            mk = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self._device)  # noqa: E731
            ws = torch.empty(int(L.bseg_workspace_bytes(self._handle, B)) + 256, dtype=torch.uint8, device=self._device)
            S = self.image_size
            st = (mk(B, 3, S, S), mk(B, 3, S, S), mk(B, 3, S, S), mk(B, 3, 2 * S, S), ws)
            self._stage[B] = st
        px, ppx, pm, pred, ws = st
        px.copy_(pixel_values.detach(), non_blocking=True)
        ppx.copy_(prompt_pixel_values.detach(), non_blocking=True)
        pm.copy_(prompt_masks.detach(), non_blocking=True)
        L = _lib.lib()
        fwd, name = ((L.bseg_forward_query_half, "bseg_forward_query_half") if query_half_only
                     else (L.bseg_forward, "bseg_forward"))
        base = (ws.data_ptr() + 255) // 256 * 256
        with torch.cuda.device(self._device):
            _lib.check(fwd(self._handle, _lib.ptr(px), _lib.ptr(ppx), _lib.ptr(pm), B,
                           0 if embedding_type == "instance" else 1, P, C.c_void_p(base),
                           C.c_size_t(ws.numel() - (base - ws.data_ptr())), _lib.ptr(pred), _lib.stream_ptr()), name)
        out = pred.clone()  # the staging buffer is overwritten by the next call
        return SegGptOutput(loss=self._hf_loss(out, labels, B, mask_rows), pred_masks=out)

    def _hf_loss(self, pred: torch.Tensor, labels: Optional[torch.Tensor], B: int, mask_rows: int = 1):
        """HF SegGptLoss (HF:modeling_seggpt.py:780-819) with the default mask == smooth-L1 over the bottom half:
        `(loss * mask).sum() / mask.sum()`.  `mask` comes from bool_masked_pos, whose DEFAULT has batch dimension 1
        (HF:modeling_seggpt.py:910-917): the numerator then runs over all B samples while the denominator counts one
        sample's 3*448*448 masked values, i.e. B times the per-sample mean -- reproduced here (`mask_rows` = the batch
        dimension of bool_masked_pos).  The reference computes this value and throws it away (src/model.py:245-255);
        kept, without a graph, for interface fidelity."""
        if labels is None:
            return None
        lab = labels.detach().to(device=self._device, dtype=torch.float32).contiguous()
        S = self.image_size
        yes = torch.ones((B, S, S), dtype=torch.uint8, device=self._device)
        loss_t = torch.empty(1, dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().bseg_loss_smoothl1_fwd_bwd(_lib.ptr(pred), _lib.ptr(lab), _lib.ptr(yes), self.beta,
                                                             1, _lib.ptr(loss_t), None, _lib.ptr(self._scratch), B,
                                                             S, S, _lib.stream_ptr()), "bseg_loss")
        # the kernel divides by B * 3*448*448 (all-ones keep mask); HF divides by mask_rows * 3*448*448
        return loss_t[0] * (float(B) / float(mask_rows)) if mask_rows != B else loss_t[0]
