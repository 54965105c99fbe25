"""Experiment configuration schema of the drop-in (same keys as the reference's YAML dict).

The key set and defaults are the ones `run_closed_loop_experiment` reads with `.get`
(reference core/tube_mpc.py:48-181, 674-786; values of configs/dubins.yaml:1-85).
"""
from __future__ import annotations

import copy
import math
from typing import Any, Dict


def default_dubins_config() -> Dict[str, Any]:
    """The Dubins + 5-obstacle experiment (reference configs/dubins.yaml) as a plain dict."""
    return copy.deepcopy({
        "seed": 0,
        "device": "cuda",
        "run_name": "dubins_strict",
        "out_dir": "outputs",
        "plot": False,
        "debug_numerics": False,
        "use_float64": True,
        "system": {
            "name": "dubins",
            "dt": 0.01,
            "horizon_N": 50,
            "task_horizon_H": 300,
            "nominal_max_iter": 10,
            "aux_max_iter": 20,
            "ilqr_reg": 1.0e-3,
            "line_search_alphas": [1.0, 0.5, 0.25, 0.1, 0.05, 0.01, 0.0],
            "control_bounds": {"v_min": -10.0, "v_max": 10.0, "omega_max": 3.141592653589793},
            "disturbance": {"w_low": [-0.05, -0.05, -0.05], "w_high": [0.05, 0.05, 0.05]},
            "target": [10.0, 10.0, 0.7853981633974483],
        },
        "dbas": {"barrier_type": "inverse", "alpha": 0.0, "gamma": 0.0, "nominal_tightening": 0.0, "eps": 1.0e-4},
        "environment": {
            "obstacles": [
                {"center": [4.0, 2.0], "radius": 1.0},
                {"center": [2.0, 4.0], "radius": 1.0},
                {"center": [4.0, 8.0], "radius": 1.0},
                {"center": [8.0, 4.0], "radius": 1.0},
                {"center": [6.0, 6.0], "radius": 1.0},
            ],
            "obstacle_smoothmin_beta": 20.0,
            "obstacle_aggregation": "smoothmin",
        },
        "cost_nominal": {"Q": [1.0, 1.0, 0.0], "R": [1.0, 1.0], "q_b": 1.0, "Qf": [1000.0, 1000.0, 1000.0]},
        "cost_auxiliary": {"Q": [1.0, 1.0, 1.0], "R": [1.0, 1.0], "q_b": 1.0},
        "adaptation": {
            "lr_eta": 5.0e-2,
            "steps": 1,
            "momentum": 0.9,
            "adapt_nominal": False,
            "adapt_ancillary": True,
            "project_params": True,
        },
    })


def load_config(path: str) -> Dict[str, Any]:
    """YAML file -> dict (same loader contract as reference run_experiment.py:30-32)."""
    import yaml

    with open(path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)


assert abs(default_dubins_config()["system"]["target"][2] - math.pi / 4) < 1e-15
