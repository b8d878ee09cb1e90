"""`engine.mcts.get_move` -- same import contract as the reference's engine/mcts/__init__.py:3-9:
if the native library is missing, get_move raises ImportError."""
try:
    from zeroclone_b200 import _ffi as _ffi
    _ffi.lib()
    from zeroclone_b200 import mcts as mcts
    get_move = mcts.get_move
except Exception:   # library not built
    mcts = None

    def get_move(*args, **kwargs):
        raise ImportError('libzc_b200.so not built (python -m zeroclone_b200.build)')
