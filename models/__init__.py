"""Drop-in import paths of the reference's `models` package (aliases of zeroclone_b200.models), so that
whole-module checkpoints pickled by the reference (scripts/train.py:143) unpickle against these classes."""
