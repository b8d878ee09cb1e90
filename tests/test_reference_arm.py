"""The reference arm of bench.py (`--impl reference`, cpu_baseline) drives the reference's STOCK files laid out by
`make -C oracle ref` into oracle/_ref/pyref.zip (unpacked to a temporary directory at run time).  CPU-only checks: the stock files are the ones imported, the root sets
the reference's own backends generate are the ones the GPU arm searches, both arms print the same `config`, and the
worker pool times searches (not start-up)."""
import hashlib
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import ref_harness as rh  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.ref_available(), reason="oracle/_ref/pyref.zip not built (needs /root/reference)")


def _stock_in_subprocess(code: str) -> str:
    """the stock `engine` package must not share a process with this repo's `engine` aliases that other tests import"""
    import subprocess
    out = subprocess.run([sys.executable, "-c", f"import sys; sys.path.insert(0, {REPO!r})\n" + code], capture_output=True, text=True,
                         timeout=300, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_pyref_files_are_the_reference_files():
    root = rh.pyref_dir()
    sums = dict(line.split()[::-1] for line in open(os.path.join(root, "SHA256SUMS")).read().splitlines())
    assert "engine/value_functions.py" in sums and "engine/games/connect4/c4_backend.py" in sums
    for rel, want in sums.items():
        got = hashlib.sha256(open(os.path.join(root, rel), "rb").read()).hexdigest()
        assert got == want, rel
        ref = os.path.join("/root/reference", rel)
        if os.path.exists(ref):           # dev container: byte-identical to the checkout
            assert hashlib.sha256(open(ref, "rb").read()).hexdigest() == want, rel


def test_root_sets_match_the_gpu_arm():
    out = _stock_in_subprocess(
        "from oracle import ref_harness as rh\n"
        "import json\n"
        "print(json.dumps({'c4': rh.c4_roots_set_b(300), 'chess': [r.hex() for r in rh.chess_roots_set_b(120)]}))\n")
    import json
    got = json.loads(out.strip().splitlines()[-1])
    from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b
    ours = c4_roots_set_b(300)
    assert [(int(r["x"]), int(r["o"]), int(r["turn"])) for r in ours] == [tuple(r) for r in got["c4"]]
    ours_ch = chess_roots_set_b(120)
    assert [r.tobytes().hex() for r in ours_ch] == got["chess"]
    # and a shard's roots are a slice of the global set (rank r owns tree ids [r*n, (r+1)*n))
    assert np.array_equal(c4_roots_set_b(50, first_tree_id=100), ours[100:150])


def test_stock_modules_and_values():
    out = _stock_in_subprocess(
        "from oracle import ref_harness as rh\n"
        "s = rh.stock(need_torch=True)\n"
        "assert s.vf.__file__.startswith(rh.pyref_dir()) and s.c4.__file__.startswith(rh.pyref_dir())\n"
        "assert s.vf.DEVICE == 'cpu'\n"
        "import sys; assert 'zeroclone_b200' not in sys.modules and not any('libzc_b200' in l for l in open('/proc/self/maps'))\n"
        "v = rh.make_value('connect4', 'c4_positional'); b = s.c4\n"
        "st = b.play_move(b.create_init_state(), (3, 0))\n"
        "assert v.batch([st], backend=b) == [-4 / 64]\n"
        "vn = rh.make_value('connect4', 'value_net')\n"
        "out = vn.batch([st, b.create_init_state()], backend=b)\n"
        "assert len(out) == 2 and all(-1 <= x <= 1 for x in out)\n"
        "mv = s.mcts.get_move(st, v, rh.first_policy, b, 64, 1.4, 32); assert mv[0] in range(7)\n"
        "print('ok')\n")
    assert out.strip().endswith("ok")


def test_pool_rate_excludes_startup():
    pool = rh.RefPool("connect4", "c4_positional", [(0, 0, 0), (1 << 21, 0, 1)], 200, cores=2)
    try:
        rate, det = pool.step(0.5)
        rate2, det2 = pool.step(0.5)
    finally:
        pool.close()
    assert det["sims"] % 200 == 0 and det["sims"] > 0
    assert 0.5 <= det["longest_worker_s"] < 1.5
    assert rate > 0 and 0.33 < rate2 / rate < 3.0          # two steps of the same pool measure the same thing (start-up is outside both)


def test_both_arms_print_the_same_config():
    import bench
    for name, wl in bench.WORKLOADS.items():
        for world in (1, 8):
            a = bench.workload_config(name, wl["trees"], wl["sims"], world)
            b = bench.workload_config(name, wl["trees"], wl["sims"], world)
            assert a == b and set(a) >= {"workload", "trees_per_gpu", "sims", "batch_size", "c", "policy", "roots", "l2"}
