"""Development aid: run the device move generator on many positions (for ncu timing of k_chess_legal_batch)."""
import ctypes as C
import sys
import numpy as np
from zeroclone_b200 import _ffi
from zeroclone_b200.workloads import chess_roots_set_b

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
base = chess_roots_set_b(4096)
arr = np.tile(base, n // 4096)
moves = np.zeros((n, _ffi.MAX_MOVES), dtype=_ffi.CHESS_MOVE_DTYPE)
counts = np.zeros(n, dtype=np.int32)
flags = np.zeros(n, dtype=np.int32)
for _ in range(2):
    _ffi.check(_ffi.lib().zc_chess_legal_moves_batch(0, arr.ctypes.data_as(C.c_void_p), n, moves.ctypes.data_as(C.c_void_p),
                                                     counts.ctypes.data_as(C.c_void_p), flags.ctypes.data_as(C.c_void_p)))
print("mean moves", counts.mean())
