"""Import the UNMODIFIED reference (/root/reference/G2Vec.py) in the build container.

TEST INFRASTRUCTURE ONLY.  The reference does `import tensorflow as tf` at module top
(G2Vec.py:3) and TensorFlow is not installed, so a stub module is registered first; only
the NumPy-only functions (steps 1-3, 5-7) are usable.  /root/reference does not exist on
the GPU box: nothing run there may call this.
"""
import importlib.util
import os
import sys
import types

REF_DIR = "/root/reference"


def available():
    return os.path.exists(os.path.join(REF_DIR, "G2Vec.py"))


def load():
    if not available():
        raise FileNotFoundError("reference not mounted at %s" % REF_DIR)
    if "tensorflow" not in sys.modules:
        sys.modules["tensorflow"] = types.ModuleType("tensorflow")
    spec = importlib.util.spec_from_file_location("g2vec_reference", os.path.join(REF_DIR, "G2Vec.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
