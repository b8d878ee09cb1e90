"""`python setup.py build_ext --inplace` in this directory (what the reference's tests run at import,
tests/test_cb.py:4, tests/test_engine_configs.py:7-8) builds the one native library of this repo."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), *[".."] * (3 if "games" in __file__ else 2))))
from zeroclone_b200.build import build  # noqa: E402

if __name__ == "__main__":
    print("built", build())
