"""Config composition for the entry points without Hydra/OmegaConf (neither is installed in this
image): a small PyYAML composer that understands exactly what the reference's tree uses —
`defaults` lists with config groups, `# @package _global_` experiment files, `${a.b}`
interpolation and `key.sub=value` command-line overrides (configs/, generate.py:203-206).
When Hydra is importable the entry points use it instead."""
import copy
import os
import re

import yaml


class Cfg(dict):
    """dict with attribute access, like the DictConfig the reference's code expects
    (`model_cfg._name_`, `cfg.pop('_name_')`, `cfg['key']`)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(x):
    if isinstance(x, dict):
        return Cfg({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)


def _load(path):
    with open(path) as f:
        text = f.read()
    return yaml.safe_load(text) or {}, bool(re.search(r"#\s*@package\s+_global_", text))


def _apply_defaults(root, out, cfg_dir, node, group_choices):
    """Process one file's `defaults` list in order; `_self_` places the file's own keys."""
    defaults = node.pop("defaults", [])
    placed_self = False
    for d in defaults:
        if d == "_self_":
            _merge(out, node)
            placed_self = True
            continue
        (group, choice), = d.items() if isinstance(d, dict) else ((d, None),)
        group = group.lstrip("/")
        choice = group_choices.get(group, choice)
        sub, is_global = _load(os.path.join(cfg_dir, group, f"{choice}.yaml"))
        target = out if is_global else out.setdefault(group, {})
        staged = {}
        _apply_defaults(root, staged, cfg_dir, sub, group_choices)
        if is_global:
            _merge(out, staged)
        else:
            _merge(target, staged)
    if not placed_self:
        _merge(out, node)


def compose(config_dir, config_name="config", overrides=()):
    """-> Cfg.  overrides: ["experiment=ljspeech", "model=wavenet", "generate.n_samples=4", ...]"""
    group_choices, assigns = {}, []
    for o in overrides:
        k, v = o.split("=", 1)
        k = k.lstrip("+")
        if "." not in k and os.path.isdir(os.path.join(config_dir, k)):
            group_choices[k] = v
        else:
            assigns.append((k, yaml.safe_load(v)))
    root, _ = _load(os.path.join(config_dir, config_name + ".yaml"))
    out = {}
    # a group chosen on the command line that the defaults tree does not mention is merged last
    _apply_defaults(root, out, config_dir, root, group_choices)
    for k, v in assigns:
        node = out
        parts = k.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v

    def resolve(x, depth=0):
        if isinstance(x, dict):
            return {k: resolve(v) for k, v in x.items()}
        if isinstance(x, list):
            return [resolve(v) for v in x]
        if isinstance(x, str):
            m = re.fullmatch(r"\$\{([^}]+)\}", x.strip())
            if m:
                node = out
                for p in m.group(1).split("."):
                    node = node[p]
                return resolve(node)
        return x

    return _wrap(resolve(out))
