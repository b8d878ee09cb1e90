"""Connect Four value network: the same tower on 2x6x7 planes (c4_backend.py:52-61).
The reference ships no Connect Four network (models/ holds only chess_value); BASELINE.json's
config 2 needs one, so it is the chess tower with a 2-plane stem."""
from ..tower import ResidualBlock, ValueNetDataset, ValueTower, safe_globals, train  # noqa: F401


class ValueNetwork(ValueTower):
    in_planes = 2


def add_safe_globals():
    safe_globals(ValueNetwork, reference_module="models.connect4_value.network")
