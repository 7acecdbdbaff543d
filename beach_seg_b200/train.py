"""The on-disk contract between the train and predict paths of the reference:
`prompt_batch.pt`, `conf.yaml`, `classes.txt` as written by src/train.py:71-77,109-122 and read back by
src/predict.py:174-178,213-216.  Files written here load in the reference (`torch.load(..., map_location="cpu")` then
`model.prompt_batch = prompt_batch`) and vice versa, so prompts trained on either engine run on the other.

prompt_batch.pt = the default-collated prompt items with "image" replaced by the list of trainable
`torch.nn.Parameter`s (src/model.py:115-130): {"image": [Parameter(3,448,448)]*N, "mask": uint8 (N,1,448,448),
"nodata": bool (N,448,448), "crop_idx": int64 (N,), "date": [str]*N, ...}, every tensor detached on the CPU
(`handle_item`, src/train.py:20-24)."""
from __future__ import annotations

import dataclasses
from pathlib import Path

import torch
import yaml

from .config import BeachSegConfig


def handle_item(v):
    """src/train.py:20-24, extended to the list of Parameters the reference leaves on the training device."""
    if isinstance(v, torch.Tensor):
        return v.detach().cpu()
    if isinstance(v, (list, tuple)) and v and isinstance(v[0], torch.Tensor):
        return [torch.nn.Parameter(x.detach().cpu().clone(), requires_grad=True) if isinstance(x, torch.nn.Parameter)
                else x.detach().cpu() for x in v]
    return v


def save_prompt_batch(model, model_dir: Path) -> Path:
    """src/train.py:76-77,121-122."""
    model_dir = Path(model_dir)
    model_dir.mkdir(parents=True, exist_ok=True)
    prompt_batch = {k: handle_item(v) for k, v in model.prompt_batch.items()}
    path = model_dir / "prompt_batch.pt"
    torch.save(prompt_batch, path)
    return path


def load_prompt_batch(model, path: Path) -> None:
    """src/predict.py:213-216 (`model.prompt_batch = torch.load(...)`), plus what the mirror needs to keep training:
    the images become the module's ParameterList on its device."""
    prompt_batch = torch.load(Path(path), map_location="cpu", weights_only=False)
    params = [torch.nn.Parameter(torch.as_tensor(p).detach().to(model.device, torch.float32), requires_grad=True)
              for p in prompt_batch["image"]]
    model.prompt_params_list = torch.nn.ParameterList(params)
    prompt_batch["image"] = params
    model.prompt_batch = prompt_batch


def _plain(v):
    if isinstance(v, Path):
        return str(v)
    if isinstance(v, (tuple, list)):
        return [_plain(x) for x in v]
    return v


def save_run_artifacts(model, conf: BeachSegConfig, model_dir: Path) -> None:
    """conf.yaml (src/train.py:111, a flat YAML mapping OmegaConf.load reads back), classes.txt (:118-119) and
    prompt_batch.pt."""
    model_dir = Path(model_dir)
    model_dir.mkdir(parents=True, exist_ok=True)
    with open(model_dir / "conf.yaml", "w") as f:
        yaml.safe_dump({k: _plain(v) for k, v in dataclasses.asdict(conf).items()}, f, sort_keys=False)
    with open(model_dir / "classes.txt", "w") as f:
        f.write("\n".join(conf.classes))
    save_prompt_batch(model, model_dir)


def load_conf(path: Path) -> BeachSegConfig:
    """src/predict.py:174-178: the training config saved next to prompt_batch.pt."""
    with open(path) as f:
        d = yaml.safe_load(f)
    fields = {f.name: f for f in dataclasses.fields(BeachSegConfig)}
    out = {}
    for k, v in d.items():
        if k not in fields:
            continue
        default = fields[k].default
        if isinstance(default, tuple) and isinstance(v, list):
            v = tuple(v)
        if isinstance(default, Path):
            v = Path(v)
        out[k] = v
    return BeachSegConfig(**out)
