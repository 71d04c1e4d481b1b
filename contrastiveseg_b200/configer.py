"""Minimal stand-in for the reference Configer (lib/utils/tools/configer.py:157-237): the loss modules
only ever call ``get(*keys)`` and ``exists(*keys)`` (lib/loss/loss_contrast.py:20-28,157-162)."""
from __future__ import annotations

import copy
import json
from typing import Any, Optional


class Configer:
    def __init__(self, config_dict: Optional[dict] = None, config_file: Optional[str] = None):
        if config_dict is None and config_file is not None:
            with open(config_file) as f:
                config_dict = json.load(f)
        self.params_root = copy.deepcopy(config_dict or {})

    def get(self, *keys) -> Any:
        v = self.params_root
        for k in keys:
            if not isinstance(v, dict) or k not in v:
                raise KeyError(f"config key {':'.join(map(str, keys))} does not exist")
            v = v[k]
        return v

    def exists(self, *keys) -> bool:
        v = self.params_root
        for k in keys:
            if not isinstance(v, dict) or k not in v:
                return False
            v = v[k]
        return True

    def add(self, keys, value) -> None:
        d = self.params_root
        for k in keys[:-1]:
            d = d.setdefault(k, {})
        d[keys[-1]] = value

    update = add

    def to_dict(self) -> dict:
        return copy.deepcopy(self.params_root)


def cityscapes_contrast_config(with_memory: bool = False) -> dict:
    """The contrast/loss/network keys of configs/cityscapes/H_48_D_4[_MEM].json:82,131,136-150."""
    d = {
        "data": {"num_classes": 19},
        "network": {"stride": 8},
        "loss": {"loss_type": "contrast_ce_loss", "params": {"ce_reduction": "elementwise_mean", "ce_ignore_index": -1}},
        "contrast": {"proj_dim": 256, "temperature": 0.1, "base_temperature": 0.07, "max_samples": 1024,
                     "max_views": 100, "stride": 8, "warmup_iters": 5000, "loss_weight": 0.1, "use_rmi": False},
    }
    if with_memory:
        d["loss"]["loss_type"] = "mem_contrast_ce_loss"
        d["contrast"].update(temperature=0.07, max_views=1, use_lovasz=False, with_memory=True, memory_size=5000,
                             pixel_update_freq=10)
    return d
