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
