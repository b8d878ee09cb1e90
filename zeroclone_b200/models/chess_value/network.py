"""Chess value network: 17x8x8 planes (chess_backend.cpp:461-521).  Module contract of the
reference's models/chess_value/network.py: ValueNetwork, ValueNetDataset, add_safe_globals, train."""
from ..tower import ResidualBlock, ValueNetDataset, ValueTower, safe_globals, train  # noqa: F401


class ValueNetwork(ValueTower):
    in_planes = 17


def add_safe_globals():
    safe_globals(ValueNetwork, reference_module="models.chess_value.network")
