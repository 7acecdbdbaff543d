"""Per-kernel-category device time of ONE forward launch at batch 1 / 4 (library's own per-launch CUDA events)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200 import _lib, synth
from beach_seg_b200.ml_util import load_model

CATS = ["gemm", "attention", "layernorm", "decoder_head", "ingest", "decode", "vote", "elementwise", "loss"]
GEMM_MODES = {0: "bf16", 1: "lin1_gelu", 2: "proj_f32(ensemble)", 3: "proj_resid", 4: "qkv", 5: "patch_embed",
              6: "dec_embed_pixshuf", 11: "lin2_resid", 14: "dec_embed_pixshuf"}
dev = torch.device("cuda:0")
model = load_model("random-init:0", device=dev, max_batch=64, graph_batch=0)
L = _lib.lib()
for B in (1, 4):
    px, ppx, pm = (t.to(dev) for t in synth.model_inputs(batch=B, seed=1))
    with torch.no_grad():
        for _ in range(3):
            model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm)
        torch.cuda.synchronize()
        L.bseg_profile_enable(1)
        reps = 10
        for _ in range(reps):
            model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm)
        torch.cuda.synchronize()
    n = len(CATS)
    pms, pl, pw, pb = (C.c_double * n)(), (C.c_longlong * n)(), (C.c_double * n)(), (C.c_double * n)()
    L.bseg_profile_collect(pms, pl, pw, pb)
    gms, gwk = (C.c_double * 16)(), (C.c_double * 16)()
    L.bseg_profile_collect_gemm(gms, gwk)
    L.bseg_profile_enable(0)
    print(f"B={B}: sum of kernel times {sum(pms) / reps:.3f} ms per forward")
    for i, c in enumerate(CATS):
        if pl[i]:
            print(f"   {c:14s} {pms[i] / reps:7.3f} ms  {pl[i] // reps:4d} launches  {pms[i] / pl[i] * 1e3:7.1f} us each"
                  + (f"  {pw[i] / (pms[i] * 1e-3) / 1e12:7.1f} TFLOP/s" if pw[i] else ""))
    for i in range(16):
        m = GEMM_MODES.get(i, str(i))
        if gms[i]:
            print(f"      gemm {m:18s} {gms[i] / reps:7.3f} ms  {gwk[i] / (gms[i] * 1e-3) / 1e12:7.1f} TFLOP/s")
