"""The reference's config surface (utils/commons/hparams.py:27-150): a module-global `hparams` dict filled by
`set_hparams()` from a yaml file with `base_config` inheritance, the saved `checkpoints/<exp>/config.yaml`
and `-hp "a=1,b.c=2,d=[1 2]"` command-line overrides.  The reference's egs/*.yaml load unchanged.

Re-implemented from the documented behaviour; the argument names, override grammar and precedence
(base configs < config < saved config unless --reset < -hp overrides) follow the reference so its CLI works.
"""
from __future__ import annotations

import argparse
import ast
import os
from typing import Any, Dict, List

import yaml

hparams: Dict[str, Any] = {}
_printed = False


def _deep_update(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_update(dst[k], v)
        else:
            dst[k] = v
    return dst


class _ConfigLoader:
    """Depth-first `base_config` resolution; a file is visited once (reference: hparams.py:62-82)."""

    def __init__(self):
        self.chain: List[str] = []
        self.seen = set()

    def load(self, path: str) -> dict:
        if not os.path.exists(path):
            return {}
        with open(path) as f:
            node = yaml.safe_load(f) or {}
        self.seen.add(path)
        merged: dict = {}
        bases = node.get("base_config", [])
        if not isinstance(bases, list):
            bases = [bases]
            node["base_config"] = bases
        for base in bases:
            if base.startswith("."):                      # relative to the including file, else to the cwd
                base = os.path.normpath(os.path.join(os.path.dirname(path), base))
            if base not in self.seen:
                _deep_update(merged, self.load(base))
        _deep_update(merged, node)
        self.chain.append(path)
        return merged


def _coerce(old: Any, text: str) -> Any:
    """Type of an override follows the existing value (reference: hparams.py:118-123)."""
    text = text.strip("'\" ")
    if text in ("True", "False") or isinstance(old, (bool, list, dict)):
        if isinstance(old, list):
            text = text.replace(" ", ",")
        return ast.literal_eval(text)
    return type(old)(text)


def apply_overrides(cfg: dict, spec: str) -> dict:
    """`a=1,b.c=2,d=[1 1 1]` — keys must already exist (except the drop-in's own `b200_*` switches), dotted keys descend into sub-dicts."""
    if not spec:
        return cfg
    depth, cur, items = 0, "", []
    for ch in spec:                                        # split on commas that are not inside [...]
        depth += ch == "["
        depth -= ch == "]"
        if ch == "," and depth == 0:
            items.append(cur)
            cur = ""
        else:
            cur += ch
    items.append(cur)
    for item in items:
        if not item.strip():
            continue
        key, value = item.split("=", 1)
        node = cfg
        parts = key.strip().split(".")
        for part in parts[:-1]:
            node = node[part]
        if parts[-1] not in node and parts[-1].startswith("b200_"):
            # the drop-in's own switches (b200_mode, b200_vocab, b200_frames, b200_train_steps, ...) are not in the reference's yaml
            # files: they may be introduced on the command line (any other unknown key is a KeyError, as in the reference)
            text = value.strip("'\" ")
            try:
                node[parts[-1]] = ast.literal_eval(text)
            except (ValueError, SyntaxError):
                node[parts[-1]] = text
            continue
        node[parts[-1]] = _coerce(node[parts[-1]], value)
    return cfg


def _parse_cli():
    ap = argparse.ArgumentParser(description="")
    ap.add_argument("--config", type=str, default="")
    ap.add_argument("--exp_name", type=str, default="")
    ap.add_argument("-hp", "--hparams", type=str, default="")
    for flag in ("infer", "validate", "reset", "remove", "debug"):
        ap.add_argument(f"--{flag}", action="store_true")
    args, unknown = ap.parse_known_args()
    if unknown:
        print("| Unknow hparams: ", unknown)
    return args


def set_hparams(config: str = "", exp_name: str = "", hparams_str: str = "", print_hparams: bool = True,
                global_hparams: bool = True) -> dict:
    global _printed
    if config == "" and exp_name == "":
        args = _parse_cli()
    else:
        args = argparse.Namespace(config=config, exp_name=exp_name, hparams=hparams_str, infer=False, validate=False,
                                  reset=False, remove=False, debug=False)
    assert args.config != "" or args.exp_name != "", "either --config or --exp_name is required"
    if args.config:
        assert os.path.exists(args.config), f"config {args.config} not found"

    work_dir = f"checkpoints/{args.exp_name}" if args.exp_name else ""
    saved_path = f"{work_dir}/config.yaml" if work_dir else ""
    saved = {}
    if saved_path and os.path.exists(saved_path):
        with open(saved_path) as f:
            saved = yaml.safe_load(f) or {}

    loader = _ConfigLoader()
    cfg: dict = {}
    if args.config:
        cfg.update(loader.load(args.config))
    if not args.reset:
        cfg.update(saved)
    cfg["work_dir"] = work_dir
    apply_overrides(cfg, args.hparams)

    if work_dir and args.remove:
        if input("REMOVE old checkpoint? Y/N [Default: N]: ").lower() == "y":
            import shutil
            shutil.rmtree(work_dir, ignore_errors=True)
    if work_dir and (not os.path.exists(saved_path) or args.reset) and not args.infer:
        os.makedirs(work_dir, exist_ok=True)
        with open(saved_path, "w") as f:
            yaml.safe_dump(cfg, f)

    cfg.update(infer=args.infer, debug=args.debug, validate=args.validate, exp_name=args.exp_name)
    if global_hparams:
        hparams.clear()
        hparams.update(cfg)
        if print_hparams and not _printed:
            print("| Hparams chains: ", loader.chain)
            print("| Hparams: ")
            for i, (k, v) in enumerate(sorted(cfg.items())):
                print(f"\033[;33;m{k}\033[0m: {v}, ", end="\n" if i % 5 == 4 else "")
            print("")
            _printed = True
    return cfg
