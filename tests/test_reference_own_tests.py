"""The reference's OWN test files, unmodified (tests/test_cb.py, tests/test_engine_configs.py -- packed into the build artefact
oracle/_ref/pyref.zip by `make -C oracle ref`), run against THIS repo: a scratch directory holds the two files under `tests/` and
symlinks `engine`, `models`, `configs` to this repo's drop-in packages, which is all those files see of their surroundings (they
insert their parent directory into sys.path and run `setup.py build_ext --inplace` in `engine/games/chess` and `engine/mcts`).
test_cb.py needs no GPU (the six backend functions over the C-ABI's host entry points); test_engine_configs.py searches.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import pytest

from oracle import ref_harness as rh

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not rh.ref_available(), reason="oracle/_ref/pyref.zip not built (needs /root/reference)")


def run_reference_test_file(name):
    src = os.path.join(rh.pyref_dir(), "tests", name)
    scratch = tempfile.mkdtemp(prefix="zc_reftests_")
    try:
        os.makedirs(os.path.join(scratch, "tests"))
        shutil.copy(src, os.path.join(scratch, "tests", name))
        for d in ("engine", "models", "configs", "zeroclone_b200"):
            os.symlink(os.path.join(REPO, d), os.path.join(scratch, d))
        env = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get("PYTHONPATH", ""))
        out = subprocess.run([sys.executable, "-m", "pytest", os.path.join("tests", name), "-q", "-p", "no:cacheprovider", "-c", os.devnull,
                              "--rootdir", scratch], cwd=scratch, env=env, capture_output=True, text=True, timeout=1200)
        return out
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


def test_reference_test_cb_passes_unmodified():
    out = run_reference_test_file("test_cb.py")
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "10 passed" in out.stdout, out.stdout[-500:]          # 5 unit tests + 5 FEN endings (test_cb.py:39-116)


@pytest.mark.gpu
def test_reference_test_engine_configs_passes_unmodified():
    out = run_reference_test_file("test_engine_configs.py")
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout, out.stdout[-500:]
