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

# BASELINE.json configs[4] "refined composite soft object (4x particle count)": the softbox scene with the composite grid
# refined from 4x5x7 at spacing 0.3 to 7x9x13 at spacing 0.15 -- the same outer box, 434 shell elements instead of 110 --
# around the reference's own gripper file (included, not copied).  Everything else as in the softbox scene
# (ref: data/gripper/soft_experiments_softbox_adjusted_for_2_fingers.xml:6-14).
REFINED = """<mujoco model="soft experiments_softbox_refined">
    <include file="{gripper}"/>
    <worldbody>
        <body pos="1.7 0 1.0">
            <composite prefix="OBJ" type="box" count="7 9 13" spacing="0.15">
                <geom type="capsule" size=".2 0.2" rgba=".8 .2 .1 1" mass="0.0005" contype="0" conaffinity="1"/>
                <joint kind="main" stiffness="700" damping="100" solreffix="-100 -10" solimpfix="0.9 0.97 0.000001 0.9 2"/>
                <tendon kind="main" stiffness="700" damping="100" solreffix="-100 -10" solimpfix="0.9 0.97 0.000001 0.9 2"/>
            </composite>
        </body>
    </worldbody>
</mujoco>
"""


def report(name, m):
    out = os.path.join(HERE, name + ".sgm")
    mjcf.save_blob(m, out)
    print(name, "nv", m.nv, "neq", m.neq, "->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    import tempfile
    for name, f in MODELS.items():
        report(name, mjcf.load_mjcf(os.path.join(REF, "data", "gripper", f)))
    with tempfile.TemporaryDirectory() as d:
        xml = os.path.join(d, "softbox_refined.xml")
        with open(xml, "w") as fh:
            fh.write(REFINED.format(gripper=os.path.join(REF, "data", "gripper", "soft_grip_two_fingers.xml")))
        report("softbox_refined", mjcf.load_mjcf(xml))
