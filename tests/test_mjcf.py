"""Model compiler: structure, MuJoCo element order, compile-time constants."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE, ROOT, blob_path

MINI = os.path.join(ROOT, "tests", "data", "mini.xml")
# (nv, nbody, ngeom, neq, nshell, mass scale, m_el)  -- derived in SURVEY App. B/D independently of this code
EXPECT = {
    "softball": (226, 229, 228, 651, 218, 0.2560018, 7.680055e-4),
    "softbox": (118, 121, 120, 327, 110, 0.3891724, 1.945862e-4),
    "softcylinder": (200, 203, 202, 573, 192, 0.2678890, 8.036671e-4),
}


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_blob_sizes_and_mass_scale(mjcf, name):
    m = mjcf.load_blob(blob_path(name))
    nv, nbody, ngeom, neq, nshell, scale, mel = EXPECT[name]
    assert (m.nv, m.nbody, m.ngeom, m.neq, m.nshell, m.ntendon) == (nv, nbody, ngeom, neq, nshell, 3)
    assert m.opt["mass_scale"] == pytest.approx(scale, rel=1e-6)
    assert m.body_mass[11] == pytest.approx(mel, rel=1e-6)
    assert m.body_mass[1:].sum() == pytest.approx(0.45, rel=1e-12)          # settotalmass
    assert m.opt["timestep"] == 0.005 and m.opt["iterations"] == 30 and m.opt["tolerance"] == 1e-7


def test_appendix_d_constants_softball(mjcf):
    """dof/body/tendon invweight0, meaninertia, tendon length: SURVEY App. D column S."""
    m = mjcf.load_blob(blob_path("softball"))
    np.testing.assert_allclose(m.dof_invweight0[:4], [33.057, 404.092, 675.176, 1112.853], rtol=1e-5)
    np.testing.assert_allclose(m.dof_invweight0[4:8], [404.092, 33.057, 675.176, 1031.889], rtol=1e-5)
    assert m.dof_invweight0[8] == pytest.approx(1302.074, rel=1e-6)
    np.testing.assert_allclose(m.body_invweight0[4], [0, 145.716], atol=1e-3)
    np.testing.assert_allclose(m.body_invweight0[5], [11.019, 450.293], rtol=1e-4)
    np.testing.assert_allclose(m.body_invweight0[9], [11.019, 423.305], rtol=1e-4)
    np.testing.assert_allclose(m.body_invweight0[11], [434.025, 0], atol=1e-3)
    np.testing.assert_allclose(m.tendon_invweight0, [283852.1, 2.5293, 2.5293], rtol=2e-5)
    np.testing.assert_allclose(m.tendon_length0, [0, 1.287012, 1.287012], atol=1e-6)
    assert m.opt["meaninertia"] == pytest.approx(1.084572e-3, rel=1e-6)


def test_appendix_d_constants_softbox(mjcf):
    m = mjcf.load_blob(blob_path("softbox"))
    np.testing.assert_allclose(m.dof_invweight0[:4], [21.745, 265.816, 444.138, 732.047], rtol=2e-5)
    assert m.dof_invweight0[8] == pytest.approx(5139.111, rel=1e-6)
    assert m.tendon_invweight0[0] == pytest.approx(565302.2, rel=1e-6)
    assert m.opt["meaninertia"] == pytest.approx(1.182249e-3, rel=1e-6)


def test_element_order_matches_mujoco(mjcf):
    """Bodies depth-first, composite tendon/equalities first, C vs C' joint order."""
    m = mjcf.load_blob(blob_path("softbox"))
    names = m.names["geom"]
    assert names[:10] == ["ground", None, None, "g121", "g122", "g123", "g21", "g22", "g23", "OBJGcenter"]
    assert names[10] == "OBJG0_0_0"
    A = m.arrays
    assert list(A["jnt_type"][:8]) == [mjcf.JNT_HINGE] * 8 and set(A["jnt_type"][8:]) == {mjcf.JNT_SLIDE}
    # left finger: hinge-z then twist-x ; right finger: twist-x then hinge-z
    np.testing.assert_array_equal(A["jnt_axis"][0], [0, 0, 1]); np.testing.assert_array_equal(A["jnt_axis"][1], [1, 0, 0])
    np.testing.assert_array_equal(A["jnt_axis"][4], [1, 0, 0]); np.testing.assert_array_equal(A["jnt_axis"][5], [0, 0, 1])
    np.testing.assert_allclose(A["jnt_range"][0], [-0.5, 0.1]); np.testing.assert_allclose(A["jnt_range"][6], [-0.4, 0.02])
    assert list(A["dof_parentid"][:9]) == [-1, 0, 1, 2, -1, 4, 5, 6, -1]
    assert A["tendon_type"][0] == mjcf.TEN_FIXED and A["tendon_num"][0] == m.nshell
    assert A["eq_type"][-1] == mjcf.EQ_TENDON and (A["eq_type"][:-1] == mjcf.EQ_JOINT).all()
    # first element: fix row then its neighbour rows, all towards later joints
    assert A["eq_obj1id"][0] == 8 and A["eq_obj2id"][0] == -1
    pair = A["eq_obj2id"][:-1] >= 0
    assert (A["eq_obj2id"][:-1][pair] > A["eq_obj1id"][:-1][pair]).all()
    assert pair.sum() == 216 and (~pair).sum() == 110


def test_blob_roundtrip(mjcf):
    m = mjcf.load_blob(blob_path("softcylinder"))
    m2 = mjcf.load_blob(mjcf.model_to_blob(m))
    for k, a in m.arrays.items():
        np.testing.assert_array_equal(a, m2.arrays[k])
    assert m2.names["geom"] == m.names["geom"] and m2.opt["nM"] == m.opt["nM"]


def test_mini_model_compiles(mjcf):
    """Own MJCF asset: defaults classes, composite expansion counts, tendon/actuator/sensor wiring."""
    m = mjcf.load_mjcf(MINI)
    assert m.nshell == 26 and m.nv == 3 + 26                    # 3x3x3 shell = 27 - 1 interior
    assert m.neq == 26 + 48 + 1                                # fix rows + neighbour rows + tendon row
    assert m.body_mass[1:].sum() == pytest.approx(0.3)
    A = m.arrays
    assert A["actuator_gain"][0] == 100 and A["actuator_trnid"][0] == 1
    assert list(A["sensor_type"]) == [mjcf.SENS_ACCEL, mjcf.SENS_GYRO]
    assert A["geom_condim"][0] == 1 and A["geom_contype"][-1] == 0
    # shell capsule axes are radial
    kin = mjcf.kinematics(m, A["qpos0"])
    centre = kin["xpos"][A["jnt_bodyid"][3] - 1]
    for j in range(3, m.nv):
        r = kin["xpos"][A["jnt_bodyid"][j]] - centre
        np.testing.assert_allclose(kin["xaxis"][j], r / np.linalg.norm(r), atol=1e-12)


def test_unsupported_elements_raise(mjcf, tmp_path):
    bad = tmp_path / "bad.xml"
    bad.write_text(open(MINI).read().replace('<body pos="1.2 0 0.6">', '<body pos="1.2 0 0.6"><freejoint/>'))
    with pytest.raises(mjcf.UnsupportedMJCF):
        mjcf.load_mjcf(str(bad))
    bad.write_text(open(MINI).read().replace('solver="PGS"', 'solver="Newton"'))
    with pytest.raises(mjcf.UnsupportedMJCF):
        mjcf.load_mjcf(str(bad))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
@pytest.mark.parametrize("name", sorted(EXPECT))
def test_committed_blobs_match_reference_xml(mjcf, name):
    """The .sgm fixtures are exactly what the compiler produces from the reference's MJCF today."""
    f = "soft_experiments_%s_adjusted_for_2_fingers.xml" % name
    m = mjcf.load_mjcf(os.path.join(REFERENCE, "data", "gripper", f))
    assert mjcf.model_to_blob(m) == open(blob_path(name), "rb").read()


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_legacy_freejoint_model_is_rejected(mjcf):
    with pytest.raises(mjcf.UnsupportedMJCF):
        mjcf.load_mjcf(os.path.join(REFERENCE, "data", "gripper", "soft_experiments_softball.xml"))
