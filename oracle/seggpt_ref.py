"""CPU oracle for the SegGPT arithmetic of the beach_seg hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
package; the product (`beach_seg_b200/`) never does.

The arithmetic of the reference path lives in a third-party dependency that is NOT vendored under the reference
tree: `transformers.models.seggpt` (reference pins no version, `environment.yml:33`; this image has 5.5.0).  Two
things live here:

  * `make_reference_model()`  – the real HF `SegGptForImageSegmentation(SegGptConfig())`, seeded random-init
    (`from_pretrained("BAAI/seggpt-vit-large")` is impossible offline).  This IS the reference arithmetic.
  * `seggpt_forward()`        – a plain-torch fp32 restatement of the same forward, line-referenced to
    `HF:modeling_seggpt.py`, which also returns intermediates so single kernels can be checked.
    It is pinned against the HF module itself in tests/test_oracle_model.py (max |diff| ~1e-6).

`HF:` = site-packages/transformers/models/seggpt/ (transformers 5.5.0).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

GRID_H, GRID_W = 56, 28
T = GRID_H * GRID_W


def make_reference_model(seed: int = 0, stress: bool = False, num_layers: int = 24, merge_index: int = 2,
                         intermediate=(5, 11, 17, 23), image_size: int = 448):
    """HF SegGPT ViT-L, seeded random init, frozen, eval (what src/util/ml_util.py:7-13 load_model returns, minus
    from_pretrained / torch.compile).

    stress=True additionally randomises every bias / LayerNorm affine / token and enlarges the rel-pos tables and
    qkv weights so that softmax rows are far from uniform: HF's default init has zero biases and std-0.02 weights,
    which would leave bias / affine / rel-pos code paths numerically untested.

    image_size != 448 builds the native-resolution variant `SegGptConfig(image_size=(2 * image_size, image_size))`
    (verified for 512 -> T = 2048 tokens; the rel-pos tables then have 2*64-1 / 2*32-1 rows).
    """
    from transformers import SegGptConfig, SegGptForImageSegmentation

    torch.manual_seed(seed)
    extra = {} if image_size == 448 else {"image_size": [2 * image_size, image_size]}
    cfg = SegGptConfig(num_hidden_layers=num_layers, merge_index=merge_index,
                       intermediate_hidden_state_indices=list(intermediate), **extra)
    model = SegGptForImageSegmentation(cfg)
    if stress:
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if name.endswith(".bias"):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.1)
                elif "layernorm" in name and name.endswith(".weight"):
                    p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)
                elif "rel_pos" in name:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.15)
                elif name.endswith("qkv.weight"):
                    p.mul_(2.5)
                elif "token" in name:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.2)
    for p in model.parameters():
        p.requires_grad_(False)
    return model.eval()


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def rel_pos_bias(q: torch.Tensor, rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, grid_h: int = GRID_H,
                 grid_w: int = GRID_W) -> torch.Tensor:
    """add_decomposed_rel_pos (HF:modeling_seggpt.py:268-311) for q_size == k_size == (grid_h, grid_w) with natively
    sized tables (2*grid-1 rows), where get_rel_pos (:232-266) is the identity resize.  q: [N, T, 64] (UNSCALED).
    Returns [N, T, T]."""
    n = q.shape[0]
    ih = torch.arange(grid_h)[:, None] - torch.arange(grid_h)[None, :] + (grid_h - 1)
    iw = torch.arange(grid_w)[:, None] - torch.arange(grid_w)[None, :] + (grid_w - 1)
    Rh = rel_pos_h[ih]  # [gh, gh, 64]
    Rw = rel_pos_w[iw]  # [gw, gw, 64]
    rq = q.reshape(n, grid_h, grid_w, -1)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    bias = rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]
    return bias.reshape(n, grid_h * grid_w, grid_h * grid_w)


def attention_ref(q, k, v, rel_pos_h, rel_pos_w, grid_h: int = GRID_H, grid_w: int = GRID_W):
    """SegGptAttention.forward core (HF:modeling_seggpt.py:324-344). q,k,v: [N, T, 64] fp32."""
    scale = q.shape[-1] ** -0.5
    attn = (q * scale) @ k.transpose(-2, -1)
    attn = attn + rel_pos_bias(q, rel_pos_h, rel_pos_w, grid_h, grid_w)
    attn = torch.softmax(attn, dim=-1, dtype=torch.float32)
    return attn @ v


def embed_table(sd: Dict[str, torch.Tensor], embedding_type: str) -> torch.Tensor:
    """Input-independent part of SegGptEmbeddings.forward (HF:modeling_seggpt.py:163-206) -> [2, T, 1024]."""
    e = "model.embeddings."
    pos = sd[e + "position_embeddings"][:, 1:]
    n = int(math.sqrt(pos.shape[1]))
    pos = F.interpolate(pos.reshape(1, n, n, -1).permute(0, 3, 1, 2), size=(GRID_H, GRID_W), mode="bicubic",
                        align_corners=False).permute(0, 2, 3, 1).reshape(T, -1)
    typ = sd[e + ("type_token_instance" if embedding_type == "instance" else "type_token_semantic")].reshape(-1)
    bias = sd[e + "patch_embeddings.projection.bias"]
    s_in = sd[e + "segment_token_input"].reshape(-1)
    s_pr = sd[e + "segment_token_prompt"].reshape(-1)
    mask_tok = sd[e + "mask_token"].reshape(-1)
    t0 = bias[None] + s_in[None] + pos + typ[None]
    base1 = bias[None].expand(T, -1).clone()
    base1[T // 2:] = mask_tok
    t1 = base1 + s_pr[None] + pos + typ[None]
    return torch.stack([t0, t1])


def seggpt_forward(sd: Dict[str, torch.Tensor], pixel_values, prompt_pixel_values, prompt_masks,
                   embedding_type: str = "instance", feature_ensemble: bool = False, num_layers: int = 24,
                   merge_index: int = 2, intermediate=(5, 11, 17, 23), eps: float = 1e-6,
                   capture: Optional[dict] = None) -> torch.Tensor:
    """Restatement of SegGptForImageSegmentation.forward (HF:modeling_seggpt.py:839-959) in eval mode with the
    default bool_masked_pos.  Returns pred_masks [B, 3, 896, 448]; `capture` (if given) receives intermediates."""
    if embedding_type not in ("instance", "semantic"):
        raise ValueError(f"Embedding type should be either 'semantic' or 'instance', but got {embedding_type}")
    B = pixel_values.shape[0]
    # HF:713-718
    img_a = torch.cat((prompt_pixel_values, pixel_values), dim=2)
    img_b = torch.cat((prompt_masks, prompt_masks), dim=2)
    # HF:108,120 patch embedding
    pw = sd["model.embeddings.patch_embeddings.projection.weight"]
    pb = sd["model.embeddings.patch_embeddings.projection.bias"]
    ea = F.conv2d(img_a, pw, pb, stride=16).permute(0, 2, 3, 1)
    eb = F.conv2d(img_b, pw, pb, stride=16).permute(0, 2, 3, 1)
    e = "model.embeddings."
    # HF:175-178 mask token on the bottom half of the prompt stream
    w = torch.cat([torch.zeros(T // 2), torch.ones(T - T // 2)]).reshape(1, GRID_H, GRID_W, 1)
    eb = eb * (1 - w) + sd[e + "mask_token"] * w
    pos = sd[e + "position_embeddings"][:, 1:]
    n = int(math.sqrt(pos.shape[1]))
    pos = F.interpolate(pos.reshape(1, n, n, -1).permute(0, 3, 1, 2), size=(GRID_H, GRID_W), mode="bicubic",
                        align_corners=False).permute(0, 2, 3, 1)
    typ = sd[e + ("type_token_instance" if embedding_type == "instance" else "type_token_semantic")]
    ea = ea + sd[e + "segment_token_input"] + pos + typ
    eb = eb + sd[e + "segment_token_prompt"] + pos + typ
    h = torch.cat((ea, eb), dim=0)  # [2B, 56, 28, 1024]
    if capture is not None:
        capture["embeddings"] = h.clone()

    inter = []
    for i in range(num_layers):
        p = f"model.encoder.layers.{i}."
        n_seq = h.shape[0]
        x = _ln(h, sd[p + "layernorm_before.weight"], sd[p + "layernorm_before.bias"], eps)
        qkv = F.linear(x, sd[p + "attention.qkv.weight"], sd[p + "attention.qkv.bias"])
        qkv = qkv.reshape(n_seq, T, 3, 16, 64).permute(2, 0, 3, 1, 4).reshape(3, n_seq * 16, T, 64)
        q, k, v = qkv.unbind(0)
        o = attention_ref(q, k, v, sd[p + "attention.rel_pos_h"], sd[p + "attention.rel_pos_w"])
        o = o.reshape(n_seq, 16, GRID_H, GRID_W, 64).permute(0, 2, 3, 1, 4).reshape(n_seq, GRID_H, GRID_W, 1024)
        if capture is not None and i == 0:
            capture["l0_ln1"] = x.clone()
            capture["l0_q"], capture["l0_k"], capture["l0_v"] = q.clone(), k.clone(), v.clone()
            capture["l0_attn"] = o.clone()
        a = F.linear(o, sd[p + "attention.proj.weight"], sd[p + "attention.proj.bias"])
        # HF:420-429 feature ensemble
        ensemble_cond = 2 if merge_index > i else 1
        if feature_ensemble and a.shape[0] // 2 >= ensemble_cond:
            prompt, inputs = a.split(a.shape[1] // 2, dim=1)
            if ensemble_cond == 2:
                num_prompts = a.shape[0] // 2
                inputs = inputs.reshape(2, num_prompts, -1)
                inputs = inputs.mean(dim=1, keepdim=True).expand_as(inputs)
                inputs = inputs.reshape(*prompt.shape)
            else:
                inputs = inputs.mean(dim=0, keepdim=True).expand_as(inputs)
            a = torch.cat([prompt, inputs], dim=1)
        h = a + h
        x = _ln(h, sd[p + "layernorm_after.weight"], sd[p + "layernorm_after.bias"], eps)
        x = F.linear(x, sd[p + "mlp.lin1.weight"], sd[p + "mlp.lin1.bias"])
        x = F.gelu(x)
        x = F.linear(x, sd[p + "mlp.lin2.weight"], sd[p + "mlp.lin2.bias"])
        h = h + x
        if i == merge_index:  # HF:476-479
            h = (h[: h.shape[0] // 2] + h[h.shape[0] // 2:]) * 0.5
        if i in intermediate:  # HF:481-482
            inter.append(_ln(h, sd["model.encoder.layernorm.weight"], sd["model.encoder.layernorm.bias"], eps))
        if capture is not None:
            capture[f"h{i}"] = h.clone()
    feats = torch.cat(inter, dim=-1)  # HF:932-933
    if capture is not None:
        capture["inter"] = feats.clone()
    # decoder HF:555-585
    x = F.linear(feats, sd["decoder.decoder_embed.weight"], sd["decoder.decoder_embed.bias"])
    x = x.reshape(B, GRID_H, GRID_W, 16, 16, 64).permute(0, 5, 1, 3, 2, 4).reshape(B, 64, GRID_H * 16, GRID_W * 16)
    if capture is not None:
        capture["dec_nchw"] = x.clone()
    x = F.conv2d(x, sd["decoder.decoder_pred.conv.weight"], sd["decoder.decoder_pred.conv.bias"], padding=1)
    x = _ln(x.permute(0, 2, 3, 1), sd["decoder.decoder_pred.layernorm.weight"],
            sd["decoder.decoder_pred.layernorm.bias"], eps).permute(0, 3, 1, 2)
    x = F.gelu(x)
    x = F.conv2d(x, sd["decoder.decoder_pred.head.weight"], sd["decoder.decoder_pred.head.bias"])
    return x
