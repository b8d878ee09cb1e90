"""Context measurement, not a bench arm: the reference's stock search with its network evaluator ON THE GPU (its Value class puts
the model on CUDA in fp16 when a GPU is visible, value_functions.py:5-6) -- how a user of the reference would run configs[1] /
configs[3] on this box.  bench.py's reference arm hides CUDA from the reference (the contract's CPU baseline).

  python tools/ref_gpu_net.py [workers] [seconds]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ref_harness as rh  # noqa: E402

if __name__ == "__main__":
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
    for game, rows, sims in (("connect4", rh.c4_roots_set_b(256), 800), ("chess", rh.chess_roots_set_b(256), 800)):
        for hide in (True, False):
            pool = rh.RefPool(game, "value_net", rows, sims, cores=workers, hide_cuda=hide)
            try:
                pool.step(2.0)
                rate, det = pool.step(secs)
            finally:
                pool.close()
            print(f"{game:9s} value net, {workers} worker processes, network on {'CPU fp32' if hide else 'GPU fp16 (stock default)'}: {rate:,.0f} sims/s ({det['trees']} trees)")
