"""Import the UNMODIFIED reference script (G2Vec.py) -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference does `import tensorflow as tf` at module top (G2Vec.py:3) and TensorFlow is not installed.
``load()`` registers oracle/tf1_shim.py as ``tensorflow`` first, so that every function of the script is
usable: the NumPy-only steps 1-3 and 5-7 as they are, and step 4 (``compute_genetovec``, :217-286) on
the shim's restatement of the TF 1.x ops it calls.

Where the script comes from: ``/root/reference/G2Vec.py`` in the build container; on the GPU box (where
/root/reference does not exist) the copy that ``stage()`` -- called by ``__graft_entry__.build()`` --
placed in the git-ignored ``oracle/_ref/`` (it travels with the snapshot the same way the built .so
files do; the reference's sources are never committed).
"""
import importlib.util
import os
import shutil
import sys

REF_DIR = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(_HERE, "_ref", "G2Vec.py")


def script_path():
    p = os.path.join(REF_DIR, "G2Vec.py")
    if os.path.exists(p):
        return p
    if os.path.exists(STAGED):
        return STAGED
    return None


def available():
    return script_path() is not None


def stage():
    """Copy the reference script, byte for byte, into oracle/_ref/ (git-ignored).  No-op without /root/reference."""
    src = os.path.join(REF_DIR, "G2Vec.py")
    if not os.path.exists(src):
        return STAGED if os.path.exists(STAGED) else None
    os.makedirs(os.path.dirname(STAGED), exist_ok=True)
    shutil.copyfile(src, STAGED)
    return STAGED


def load():
    path = script_path()
    if path is None:
        raise FileNotFoundError("reference script not found (neither %s nor %s)" % (REF_DIR, STAGED))
    from . import tf1_shim
    tf1_shim.install()
    spec = importlib.util.spec_from_file_location("g2vec_reference", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
