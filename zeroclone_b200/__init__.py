"""zeroclone_b200 -- B200-native batched MCTS self-play engine behind the ZeroClone Engine API."""
__version__ = "0.1.0"
