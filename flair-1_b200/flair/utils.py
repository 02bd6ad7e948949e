"""YAML reading and the config recap printer (mirrors src/flair/utils.py)."""
from __future__ import annotations

import yaml


def read_config(file_path: str) -> dict:
    """src/flair/utils.py:5-7."""
    with open(file_path, "r") as f:
        return yaml.safe_load(f)


def print_recap(config: dict, dict_train, dict_val, dict_test) -> None:
    """src/flair/utils.py:10-34: indented dump of the config tree and the sample counts."""
    def walk_config(cfg, prefix=""):
        for k, v in cfg.items():
            if isinstance(v, dict):
                print(f"{prefix}|- {k}:")
                walk_config(v, prefix=prefix + "|   ")
            elif isinstance(v, list):
                print(f"{prefix}|- {k}: {v}")
            else:
                print(f"{prefix}|- {k}: {v}")
    print("Configuration Tree:")
    walk_config(config, "")
    list_keys = ["channels", "classes", "norm_means", "norm_stds"]
    for k in list_keys:
        if k in config:
            pass
    print("")
    print("[---DATA SPLIT---]")
    for name, d in (("train", dict_train), ("val", dict_val), ("test", dict_test)):
        if d:
            print(f"- {name:11s}", "", len(d["IMG"]), "samples")
    print("")
