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


# buffers of GaussianDiffusion whose LENGTH is timesteps + 1 (spec_denoiser.py:47-69): a checkpoint trained with timesteps = 8 cannot be
# strict-loaded into a 100-step model; they are functions of `timesteps` alone and are rebuilt by the module.
SCHEDULE_BUFFERS = ("timesteps", "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                    "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
                    "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2")


def load_ckpt(cur_model, ckpt_base_dir, model_name="model", force=True, strict=True, drop_keys=()):
    """Same call and behaviour as the reference's load_ckpt (utils/commons/ckpt_utils.py:26-66): newest checkpoint of a directory (or a
    file) into `cur_model`; strict by default; a missing checkpoint is an error when `force`; in non-strict mode shape-mismatched keys
    are printed and dropped, and the keys `load_state_dict` reports as missing / unexpected are printed too (never silent).
    `drop_keys` (an addition): names removed from the checkpoint before loading — pass SCHEDULE_BUFFERS to load a checkpoint trained with
    another `timesteps` strictly in everything else."""
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
    sd = dict(extract_state_dict(checkpoint, model_name))
    own = cur_model.state_dict()
    dropped = [k for k in drop_keys if k in sd]
    for k in dropped:
        del sd[k]
    if not strict:
        for k in [k for k, v in sd.items() if k in own and own[k].shape != v.shape]:
            print("| Unmatched keys: ", k, tuple(own[k].shape), tuple(sd[k].shape))
            del sd[k]
        res = cur_model.load_state_dict(sd, strict=False)
        if res.missing_keys:
            print("| Missing keys (left at their initial values): ", list(res.missing_keys))
        if res.unexpected_keys:
            print("| Unexpected keys (ignored): ", list(res.unexpected_keys))
    else:
        res = cur_model.load_state_dict(sd, strict=False)
        missing = [k for k in res.missing_keys if k not in dropped]
        if missing or res.unexpected_keys:
            raise RuntimeError(f"Error(s) in loading state_dict for {type(cur_model).__name__} from '{ckpt_path}': missing keys {missing}, "
                               f"unexpected keys {list(res.unexpected_keys)}")
    print(f"| load '{model_name}' from '{ckpt_path}'." + (f" (rebuilt from `timesteps`, not loaded: {dropped})" if dropped else ""))
    return ckpt_path


def save_ckpt(model, work_dir, global_step, optimizer=None, epoch=0, best_val=None, num_ckpt_keep=3, model_name="model"):
    """Write `<work_dir>/model_ckpt_steps_<global_step>.ckpt` in the layout the reference's Trainer writes and its `load_ckpt` /
    `restore_weights` / `restore_opt_state` read (utils/commons/trainer.py:431-470): {'state_dict': {model_name: sd}, 'global_step',
    'epoch', 'checkpoint_callback_best', 'optimizer_states'}; atomically (`.part` then rename, :451-455), then all but the newest
    `num_ckpt_keep` checkpoints are removed (:436-438).  Returns the path."""
    os.makedirs(work_dir, exist_ok=True)
    path = f"{work_dir}/model_ckpt_steps_{int(global_step)}.ckpt"
    checkpoint = {"epoch": int(epoch), "global_step": int(global_step), "checkpoint_callback_best": best_val,
                  "optimizer_states": [optimizer.state_dict()] if optimizer is not None else [],
                  "state_dict": {model_name: {k: v.detach().cpu() for k, v in model.state_dict().items()}}}
    tmp = path + ".part"
    torch.save(checkpoint, tmp)
    os.replace(tmp, path)
    for old in get_all_ckpts(work_dir)[int(num_ckpt_keep):]:
        os.remove(old)
    return path


def resume(model, work_dir, optimizer=None, model_name="model", strict=True, drop_keys=()):
    """The trainer's resume (utils/commons/trainer.py:372-429: restore_weights + restore_opt_state) for one model / one optimizer:
    loads the newest checkpoint of `work_dir` when there is one and returns its (global_step, epoch), else (0, 0)."""
    checkpoint, path = get_last_checkpoint(work_dir)
    if checkpoint is None:
        return 0, 0
    load_ckpt(model, path, model_name, force=True, strict=strict, drop_keys=drop_keys)
    states = checkpoint.get("optimizer_states") or []
    if optimizer is not None and states:
        optimizer.load_state_dict(states[0])
        dev = next(model.parameters()).device
        for state in optimizer.state.values():              # optimizer state follows the parameters' device (trainer.py:414-421)
            for k, v in state.items():
                if isinstance(v, torch.Tensor):
                    state[k] = v.to(dev) if v.dim() > 0 or k != "step" else v
    return int(checkpoint.get("global_step", 0)), int(checkpoint.get("epoch", 0))
