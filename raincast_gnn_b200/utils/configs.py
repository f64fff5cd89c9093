"""The twelve run configurations the reference ships as trained_models/<leadtime>_<kind>/params.json
(SURVEY.md 2): {24h, 72h, 120h} x {normal, normal_mixed, mixed, mixed_u}.  `heads` is unused by the model;
`grad_u` is the STRING "True"/"False" (models/gnn.py:98).  write_params_json() materialises one as a run dir."""
from __future__ import annotations

import json
import os

_COMMON = {"batch_size": 8, "gnn_hidden": 128, "gnn_layers": 4, "heads": 8, "lr": 0.0001, "max_dist": 100,
           "max_epochs": 20, "u": 1.71, "xi": 0.5}
_KINDS = {"normal": ("NormalCRPS", "False"), "normal_mixed": ("MixedNormalCRPS", "False"),
          "mixed": ("MixedLoss", "False"), "mixed_u": ("MixedLoss", "True")}


def reference_config(leadtime: str = "24h", kind: str = "mixed_u") -> dict:
    loss, grad_u = _KINDS[kind]
    cfg = dict(_COMMON, loss=loss, grad_u=grad_u)
    if leadtime == "24h" and kind == "normal_mixed":
        cfg["max_dist"] = 1            # the one shipped config whose graph has self loops only
    return cfg


def write_params_json(run_dir: str, leadtime: str = "24h", kind: str = "mixed_u", **overrides) -> str:
    os.makedirs(run_dir, exist_ok=True)
    path = os.path.join(run_dir, "params.json")
    with open(path, "w") as f:
        json.dump(dict(reference_config(leadtime, kind), **overrides), f, indent=1)
    return path
