"""Compiles the reference's MJCF models into the .sgm blobs committed next to this script.

Runs only where /root/reference exists (the build container); the GPU box uses the committed blobs.
    python tests/golden/make_models.py
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
mjcf = importlib.import_module("soft-grip_b200.mjcf")
REF = os.environ.get("SOFTGRIP_REFERENCE", "/root/reference")
MODELS = {"softbox": "soft_experiments_softbox_adjusted_for_2_fingers.xml",
          "softball": "soft_experiments_softball_adjusted_for_2_fingers.xml",
          "softcylinder": "soft_experiments_softcylinder_adjusted_for_2_fingers.xml"}

if __name__ == "__main__":
    for name, f in MODELS.items():
        m = mjcf.load_mjcf(os.path.join(REF, "data", "gripper", f))
        out = os.path.join(HERE, name + ".sgm")
        mjcf.save_blob(m, out)
        print(name, "nv", m.nv, "neq", m.neq, "->", out, os.path.getsize(out), "bytes")
