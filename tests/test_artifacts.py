"""CPU: the on-disk contract between train and predict (prompt_batch.pt / conf.yaml / classes.txt,
src/train.py:71-77,109-122; src/predict.py:174-178,213-216)."""
import dataclasses

import torch

from beach_seg_b200 import train as artifacts
from beach_seg_b200.config import BeachSegConfig


class _Model:
    """The two attributes the artifact functions touch (a PromptModel needs a GPU to construct)."""

    device = torch.device("cpu")

    def __init__(self):
        from torch.utils.data import default_collate

        items = [{"image": torch.rand(3, 448, 448), "mask": torch.randint(0, 4, (1, 448, 448), dtype=torch.uint8),
                  "nodata": torch.zeros(448, 448, dtype=torch.bool), "crop_idx": i, "date": f"2024010{i}"}
                 for i in range(3)]
        self.prompt_batch = default_collate(items)
        params = [torch.nn.Parameter(self.prompt_batch["image"][i].clone()) for i in range(3)]
        self.prompt_params_list = torch.nn.ParameterList(params)
        self.prompt_batch["image"] = params


def test_prompt_batch_round_trip_and_reference_layout(tmp_path):
    m = _Model()
    conf = BeachSegConfig(epochs=3, lr=2e-3)
    artifacts.save_run_artifacts(m, conf, tmp_path)
    # what src/predict.py:213-216 does
    pb = torch.load(tmp_path / "prompt_batch.pt", map_location="cpu", weights_only=False)
    assert set(pb) == {"image", "mask", "nodata", "crop_idx", "date"}
    assert isinstance(pb["image"], list) and len(pb["image"]) == 3
    assert all(isinstance(p, torch.nn.Parameter) and p.shape == (3, 448, 448) and p.device.type == "cpu"
               for p in pb["image"])
    assert pb["mask"].shape == (3, 1, 448, 448) and pb["mask"].dtype == torch.uint8
    assert pb["crop_idx"].tolist() == [0, 1, 2] and pb["date"] == ["20240100", "20240101", "20240102"]
    # the gather the reference's prepare_prompt performs on every key (src/model.py:189-192)
    picked = {k: [v[i] for i in [2, 0]] for k, v in pb.items()}
    assert torch.equal(picked["image"][0], m.prompt_batch["image"][2].detach())
    # load back into a module and keep training
    m2 = _Model()
    artifacts.load_prompt_batch(m2, tmp_path / "prompt_batch.pt")
    assert all(torch.equal(a.detach(), b.detach()) for a, b in zip(m2.prompt_params_list, m.prompt_params_list))
    assert all(p.requires_grad for p in m2.prompt_params_list)
    # conf.yaml / classes.txt
    conf2 = artifacts.load_conf(tmp_path / "conf.yaml")
    assert dataclasses.asdict(conf2) == dataclasses.asdict(conf)
    assert (tmp_path / "classes.txt").read_text().split("\n") == list(conf.classes)


def _cpu_prompt_model(conf, n_prompts=1):
    """A PromptModel on the CPU for host-only logic (schedules, folders): the backbone is never called."""
    from beach_seg_b200.model import PromptModel

    class _NoBackbone:
        device = torch.device("cpu")

    m = PromptModel(conf, device="cpu", model=_NoBackbone())
    m.prompt_params_list = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(3)) for _ in range(n_prompts)])
    return m


def test_lr_schedule_matches_reference_configure_optimizers(golden_dir):
    """AdamW + optional linear warm-up + CosineAnnealingLR per epoch against trajectories produced by the reference's
    own `PromptModel.configure_optimizers` (src/model.py:385-428; oracle/make_golden_train.py)."""
    import numpy as np

    from oracle.make_golden_train import LR_CASES, LR_EPOCHS

    g = np.load(golden_dir / "train_golden.npz")
    for name, kw in LR_CASES.items():
        m = _cpu_prompt_model(BeachSegConfig(**kw))
        cfg = m.configure_optimizers()
        opt, sched = cfg["optimizer"], cfg["lr_scheduler"]["scheduler"]
        assert isinstance(opt, torch.optim.AdamW)
        assert cfg["lr_scheduler"]["interval"] == "epoch" and cfg["lr_scheduler"]["frequency"] == 1
        got = []
        for _ in range(LR_EPOCHS):
            got.append(opt.param_groups[0]["lr"])
            opt.step()
            sched.step()
        np.testing.assert_allclose(np.array(got), g[f"lr_{name}"], rtol=1e-12, atol=0, err_msg=name)


def test_reference_written_prompt_batch_layout(golden_dir, tmp_path):
    """tests/golden/ref_prompt_batch.pt.gz was written by the reference's own create_trainable_params + handle_item +
    torch.save (oracle/make_golden_train.py).  `load_prompt_batch` must accept it and the file this engine writes must
    have the same keys, types and values."""
    import gzip

    path = tmp_path / "prompt_batch.pt"
    path.write_bytes(gzip.open(golden_dir / "ref_prompt_batch.pt.gz").read())
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert set(raw) == {"image", "mask", "nodata", "crop_idx", "date"}
    assert isinstance(raw["image"], list) and all(isinstance(p, torch.nn.Parameter) for p in raw["image"])
    m = _Model()
    artifacts.load_prompt_batch(m, path)
    assert len(m.prompt_params_list) == 2 and m.prompt_batch["date"] == raw["date"]
    assert torch.equal(m.prompt_batch["mask"], raw["mask"]) and torch.equal(m.prompt_batch["crop_idx"], raw["crop_idx"])
    artifacts.save_prompt_batch(m, tmp_path / "ours")
    ours = torch.load(tmp_path / "ours" / "prompt_batch.pt", map_location="cpu", weights_only=False)
    assert set(ours) == set(raw)
    for k in raw:
        a, b = ours[k], raw[k]
        assert type(a) is type(b), k
        if isinstance(b, torch.Tensor):
            assert a.dtype == b.dtype and torch.equal(a, b), k
        elif k == "image":
            assert all(type(x) is type(y) and torch.equal(x.detach(), y.detach()) and x.requires_grad == y.requires_grad
                       for x, y in zip(a, b))
        else:
            assert a == b, k


def test_accumulator_output_folders_match_the_reference(tmp_path):
    """src/predict.py:70-75,105-112: overlays in save_dir/images, class PNGs in save_dir/masks, GeoTIFFs in save_dir/tif;
    leaving the context without a single update asserts, as the reference's save_current does."""
    import pytest

    from beach_seg_b200.predict import Accumulator

    acc = Accumulator((8, 8), tmp_path, device="cpu")
    assert sorted(p.name for p in tmp_path.iterdir()) == ["images", "masks", "tif"]
    assert (acc.img_dir, acc.mask_dir, acc.tif_dir) == (tmp_path / "images", tmp_path / "masks", tmp_path / "tif")
    with pytest.raises(AssertionError):
        with acc:
            pass


def test_fit_default_length_mirrors_trainer_max_epochs():
    """src/train.py:98: max_epochs = conf.epochs * len(prompt_batch) (number of KEYS of the saved dict); the cosine
    schedule keeps T_max = conf.epochs (src/model.py:417)."""
    conf = BeachSegConfig(epochs=2)
    m = _cpu_prompt_model(conf, n_prompts=3)
    m.prompt_batch = {"image": list(m.prompt_params_list), "mask": None, "nodata": None, "crop_idx": None, "date": None}
    steps = []
    m.training_step = lambda batch, i: (m.prompt_params_list[0] ** 2).sum()  # host-only stand-in for the CUDA step
    m.fit([{"x": 0}], on_step=lambda i, loss: steps.append(i))
    assert len(steps) == conf.epochs * 5
