import importlib
import sys


def alias(name: str, target: str) -> None:
    """Make `name` in sys.modules the very module object `target` (same attributes, monkeypatchable)."""
    sys.modules[name] = importlib.import_module(target)
