"""CPU: pins oracle/seggpt_ref.py.  (i) the seeded HF module reproduces the committed golden slice (so the GPU box
regenerates the same reference weights); (ii) the plain-torch restatement equals the HF module."""
import numpy as np
import pytest
import torch

from beach_seg_b200 import synth
from oracle.seggpt_ref import make_reference_model, seggpt_forward

SMALL = dict(num_layers=5, merge_index=1, intermediate=(1, 2, 3, 4))


def test_hf_module_matches_golden_slice(golden_dir):
    g = np.load(golden_dir / "seggpt_golden.npz")
    model = make_reference_model(seed=0, stress=True)
    px, ppx, pm = synth.model_inputs(batch=1, seed=123)
    with torch.no_grad():
        pred = model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, embedding_type="instance").pred_masks
    assert pred.shape == (1, 3, 896, 448)
    np.testing.assert_allclose(pred[:, :, ::16, ::16].numpy(), g["stress_slice"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(pred[0, :, 500, :].numpy(), g["stress_row500"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("embedding_type", ["instance", "semantic"])
def test_restatement_equals_hf_small(embedding_type):
    model = make_reference_model(seed=1, stress=True, **SMALL)
    sd = model.state_dict()
    px, ppx, pm = synth.model_inputs(batch=1, seed=9)
    with torch.no_grad():
        want = model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm,
                     embedding_type=embedding_type).pred_masks
        got = seggpt_forward(sd, px, ppx, pm, embedding_type=embedding_type, **SMALL)
    assert (got - want).abs().max().item() < 1e-4 * max(1.0, want.abs().max().item())


def test_restatement_feature_ensemble():
    model = make_reference_model(seed=2, stress=True, **SMALL)
    sd = model.state_dict()
    px, ppx, pm = synth.model_inputs(batch=2, seed=10)
    px = px[:1].expand(2, -1, -1, -1).contiguous()  # same query image, two prompts (src/predict_no_prompt.py:283-295)
    with torch.no_grad():
        want = model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, embedding_type="instance",
                     feature_ensemble=True).pred_masks
        got = seggpt_forward(sd, px, ppx, pm, feature_ensemble=True, **SMALL)
        plain = seggpt_forward(sd, px, ppx, pm, feature_ensemble=False, **SMALL)
    assert (got - want).abs().max().item() < 1e-4 * max(1.0, want.abs().max().item())
    assert (plain - want).abs().max().item() > 1e-3  # the ensemble path really changes the result


def test_bad_embedding_type_raises():
    with pytest.raises(ValueError):
        seggpt_forward({}, torch.zeros(1, 3, 448, 448), torch.zeros(1, 3, 448, 448), torch.zeros(1, 3, 448, 448),
                       embedding_type="panoptic")
