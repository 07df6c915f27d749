"""Checkpoint discovery / loading with the reference's on-disk layout (utils/commons/ckpt_utils.py:7-66,
utils/commons/trainer.py:457-470):  <dir>/model_ckpt_steps_<N>.ckpt = {'state_dict': {<model_name>: sd}, ...}."""
from __future__ import annotations

import glob
import os
import re

import torch


def get_last_checkpoint(work_dir, steps=None):
    paths = get_all_ckpts(work_dir, steps)
    if not paths:
        return None, None
    return torch.load(paths[0], map_location="cpu", weights_only=False), paths[0]


def get_all_ckpts(work_dir, steps=None):
    pattern = f"{work_dir}/model_ckpt_steps_*.ckpt" if steps is None else f"{work_dir}/model_ckpt_steps_{steps}.ckpt"
    def step_of(p):
        m = re.findall(r".*steps\_(\d+)\.ckpt", p)
        return int(m[0]) if m else -1
    return sorted(glob.glob(pattern), key=lambda p: -step_of(p))


def extract_state_dict(checkpoint: dict, model_name: str = "model") -> dict:
    sd = checkpoint["state_dict"]
    if any(k.startswith("model.") for k in sd) and model_name not in sd:     # flat lightning-style checkpoint
        return {k[len(model_name) + 1:]: v for k, v in sd.items() if k.startswith(f"{model_name}.")}
    if "." not in model_name:
        return sd[model_name]
    base, rest = model_name.split(".", 1)
    return {k[len(rest) + 1:]: v for k, v in sd[base].items() if k.startswith(f"{rest}.")}


def load_ckpt(cur_model, ckpt_base_dir, model_name="model", force=True, strict=True):
    """Same call as the reference's load_ckpt: newest checkpoint of a directory (or a file) into `cur_model`."""
    if os.path.isfile(ckpt_base_dir):
        ckpt_path, checkpoint = ckpt_base_dir, torch.load(ckpt_base_dir, map_location="cpu", weights_only=False)
    else:
        checkpoint, ckpt_path = get_last_checkpoint(ckpt_base_dir)
    if checkpoint is None:
        msg = f"| ckpt not found in {ckpt_base_dir}."
        if force:
            raise FileNotFoundError(msg)
        print(msg)
        return None
    sd = extract_state_dict(checkpoint, model_name)
    if not strict:
        own = cur_model.state_dict()
        sd = {k: v for k, v in sd.items() if k in own and own[k].shape == v.shape}
    cur_model.load_state_dict(sd, strict=strict)
    print(f"| load '{model_name}' from '{ckpt_path}'.")
    return ckpt_path
