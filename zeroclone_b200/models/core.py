"""Where the value networks live: `<this dir>/<model_type>/network.py` defines the module,
`latest.pth` beside it is the current checkpoint and `checkpoints/` its rotated predecessors --
the contract of the reference's models/core.py:10-20."""
from __future__ import annotations

import importlib
from pathlib import Path
from typing import List, Tuple

_ROOT = Path(__file__).resolve().parent


def _home(model_type: str) -> Path:
    return _ROOT / model_type


def get_value_network(model_type: str) -> Tuple[object, Path]:
    """(python module of the network, path of its latest checkpoint -- which need not exist yet)"""
    return importlib.import_module(f"{__package__}.{model_type}.network"), _home(model_type) / "latest.pth"


def list_checkpoints(model_type: str) -> List[Path]:
    """rotated checkpoints, oldest first (their names are timestamps)"""
    return sorted(_home(model_type).glob("checkpoints/*.pth"))
