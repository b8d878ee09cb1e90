"""Value-network registry: same contract as the reference's models/core.py:10-20
(`get_value_network(model_type) -> (module, latest_path)`, `list_checkpoints`)."""
from __future__ import annotations

import importlib
from pathlib import Path
from typing import List, Tuple

_ROOT = Path(__file__).resolve().parent


def get_value_network(model_type: str) -> Tuple[object, Path]:
    module = importlib.import_module(f"{__package__}.{model_type}.network")
    return module, _ROOT / model_type / "latest.pth"


def list_checkpoints(model_type: str) -> List[Path]:
    ckpt = _ROOT / model_type / "checkpoints"
    return sorted(ckpt.glob("*.pth")) if ckpt.exists() else []
