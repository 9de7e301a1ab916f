"""MJCF-subset model compiler: XML -> flat constant tables ("mjModel"-like) -> binary blob.

This is the host-side replacement for what the reference obtains from
``mujoco_py.load_model_from_path`` (ref: environment/manenv.py:27,36), restricted to exactly the
MJCF subset used by the reference's ``data/gripper`` models:

* ``<include>``, ``<compiler angle inertiafromgeom settotalmass balanceinertia>``,
  ``<option timestep solver iterations tolerance cone impratio gravity>``, ``<size>``,
  nested ``<default class>`` trees (geom / joint / site / tendon children);
* ``<body pos quat>``, ``<geom type=plane|box|capsule|sphere ...>``, ``<joint type=hinge|slide>``,
  ``<site>``;
* legacy ``<composite type=box|ellipsoid>`` shells (centre sphere, slide joints, "fix" and
  neighbour joint equalities, the fixed "volume" tendon and its tendon equality);
* ``<tendon><spatial>`` with two sites, ``<actuator><cylinder tendon area>``,
  ``<sensor><accelerometer|gyro site>``.

Anything else (``<freejoint>``, meshes, other orientation specifiers, ...) raises
``UnsupportedMJCF`` -- there is no silent approximation.

The compile rules follow SURVEY.md App. A0 (recalled MuJoCo 2.x ``user_model.cc`` /
``user_composite.cc`` / ``engine_setconst.c`` semantics; unverifiable offline because MuJoCo is
not installable here -- "parity unpinned").  Element ids come out in MuJoCo's order: bodies in
depth-first pre-order, joints/geoms/sites grouped per body in that order, the composite's tendon and
equalities before anything in the ``<tendon>``/``<equality>`` sections.
"""
from __future__ import annotations

import math
import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

__all__ = ["UnsupportedMJCF", "Model", "load_mjcf", "save_blob", "load_blob", "model_to_blob",
           "GEOM_PLANE", "GEOM_SPHERE", "GEOM_CAPSULE", "GEOM_BOX", "JNT_SLIDE", "JNT_HINGE",
           "EQ_JOINT", "EQ_TENDON", "TEN_FIXED", "TEN_SPATIAL", "SENS_ACCEL", "SENS_GYRO"]

# type codes (MuJoCo's numeric values, so contact ordering by geom type matches: plane<sphere<capsule<box)
GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX = 0, 2, 3, 6
JNT_SLIDE, JNT_HINGE = 2, 3
EQ_JOINT, EQ_TENDON = 2, 3
TEN_FIXED, TEN_SPATIAL = 0, 1
SENS_ACCEL, SENS_GYRO = 0, 1

MINVAL = 1e-15


class UnsupportedMJCF(ValueError):
    """The file uses an MJCF feature outside the subset this compiler restates."""


# --------------------------------------------------------------------------------------------
# small quaternion / rotation helpers (w, x, y, z)
# --------------------------------------------------------------------------------------------
def _normalize(v):
    v = np.asarray(v, dtype=np.float64)
    n = np.linalg.norm(v)
    return v / n if n > 0 else v


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def axis_angle_quat(axis, angle):
    s = math.sin(0.5 * angle)
    return np.array([math.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def z2quat(vec):
    """Quaternion rotating +z onto ``vec`` (MuJoCo ``mjuu_z2quat``)."""
    vec = _normalize(vec)
    axis = np.cross([0.0, 0.0, 1.0], vec)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1.0, 0.0, 0.0])
    else:
        axis = axis / s
    ang = math.atan2(s, vec[2])
    return axis_angle_quat(axis, ang)


# --------------------------------------------------------------------------------------------
# model container
# --------------------------------------------------------------------------------------------
@dataclass
class Model:
    """Flat tables. Every numpy field is serialised into the blob under its own name."""
    name: str = ""
    # options
    opt: dict = field(default_factory=dict)          # timestep, gravity, iterations, tolerance, impratio
    arrays: dict = field(default_factory=dict)       # name -> np.ndarray
    names: dict = field(default_factory=dict)        # "body"/"geom"/"joint"/"site"/"tendon" -> list[str|None]

    def __getattr__(self, k):
        arrays = object.__getattribute__(self, "arrays")
        if k in arrays:
            return arrays[k]
        raise AttributeError(k)

    @property
    def nv(self):
        return int(self.arrays["dof_bodyid"].shape[0])

    @property
    def nbody(self):
        return int(self.arrays["body_parentid"].shape[0])

    @property
    def ngeom(self):
        return int(self.arrays["geom_type"].shape[0])

    @property
    def neq(self):
        return int(self.arrays["eq_type"].shape[0])

    @property
    def ntendon(self):
        return int(self.arrays["tendon_type"].shape[0])

    @property
    def nshell(self):
        return int((self.arrays["jnt_type"] == JNT_SLIDE).sum())

    def geom_id2name(self, gid):
        return self.names["geom"][gid]


# --------------------------------------------------------------------------------------------
# XML loading with <include> splicing
# --------------------------------------------------------------------------------------------
def _load_spliced(path, _seen=None):
    _seen = set() if _seen is None else _seen
    apath = os.path.abspath(path)
    if apath in _seen:
        raise UnsupportedMJCF("recursive include of %s" % path)
    _seen.add(apath)
    root = ET.parse(apath).getroot()
    if root.tag != "mujoco":
        raise UnsupportedMJCF("root element must be <mujoco> in %s" % path)
    _splice(root, os.path.dirname(apath), _seen)
    return root


def _splice(elem, base, seen):
    out = []
    for ch in list(elem):
        if ch.tag == "include":
            inc = _load_spliced(os.path.join(base, ch.attrib["file"]), seen)
            out.extend(list(inc))
        else:
            _splice(ch, base, seen)
            out.append(ch)
    elem[:] = out


def _floats(s, n=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.shape[0] != n:
        raise UnsupportedMJCF("expected %d numbers, got %r" % (n, s))
    return v


def _bool(s):
    return str(s).strip().lower() == "true"


# --------------------------------------------------------------------------------------------
# defaults
# --------------------------------------------------------------------------------------------
_DEF_TAGS = ("geom", "joint", "site", "tendon", "general", "cylinder")


def _collect_defaults(root):
    classes = {"main": {t: {} for t in _DEF_TAGS}}

    def rec(delem, parent_name, name):
        cur = {t: dict(classes[parent_name][t]) for t in _DEF_TAGS} if parent_name else classes["main"]
        for ch in delem:
            if ch.tag == "default":
                continue
            if ch.tag in _DEF_TAGS:
                cur[ch.tag].update(ch.attrib)
            elif ch.tag in ("mesh", "material", "light", "camera", "pair", "equality", "motor",
                            "position", "velocity"):
                pass  # irrelevant to the supported physics subset
        classes[name] = cur
        for ch in delem:
            if ch.tag == "default":
                cname = ch.attrib.get("class")
                if cname is None:
                    raise UnsupportedMJCF("nested <default> needs a class name")
                rec(ch, name, cname)

    for d in root.findall("default"):
        rec(d, None if "class" not in d.attrib else "main", d.attrib.get("class", "main"))
    return classes


# --------------------------------------------------------------------------------------------
# builder
# --------------------------------------------------------------------------------------------
class _Builder:
    def __init__(self, root):
        self.root = root
        self.defaults = _collect_defaults(root)
        comp = {}
        for c in root.findall("compiler"):
            comp.update(c.attrib)
        self.angle_scale = 1.0 if comp.get("angle", "degree") == "radian" else math.pi / 180.0
        self.settotalmass = float(comp.get("settotalmass", -1))
        if comp.get("inertiafromgeom", "auto") not in ("auto", "true"):
            raise UnsupportedMJCF("inertiafromgeom must be auto/true")
        if comp.get("coordinate", "local") != "local":
            raise UnsupportedMJCF("global coordinates")
        opt = {"timestep": 0.002, "gravity": "0 0 -9.81", "iterations": 100, "tolerance": 1e-8,
               "impratio": 1.0, "solver": "Newton", "cone": "pyramidal", "integrator": "Euler"}
        for o in root.findall("option"):
            opt.update(o.attrib)
        if opt["solver"] != "PGS" or opt["cone"] != "elliptic" or opt["integrator"] != "Euler":
            raise UnsupportedMJCF("only solver=PGS cone=elliptic integrator=Euler is restated "
                                  "(ref: data/gripper/soft_scene.xml:13)")
        self.opt = {"timestep": float(opt["timestep"]), "gravity": _floats(opt["gravity"], 3),
                    "iterations": int(opt["iterations"]), "tolerance": float(opt["tolerance"]),
                    "impratio": float(opt["impratio"])}
        size = {}
        for s in root.findall("size"):
            size.update(s.attrib)
        self.opt["nconmax"] = int(size.get("nconmax", 100))
        self.opt["njmax"] = int(size.get("njmax", 500))
        # element lists
        self.bodies = []   # dict(name,parent,pos,quat,geoms[],joints[],sites[],children[])
        self.tendons = []  # dict(name,type,stiffness,damping,wraps)
        self.equalities = []
        self.actuators = []
        self.sensors = []

    # ---- attribute resolution ----
    def _attrs(self, elem, tag, childclass):
        cls = elem.attrib.get("class", childclass or "main")
        if cls not in self.defaults:
            raise UnsupportedMJCF("unknown default class %r" % cls)
        a = dict(self.defaults[cls][tag])
        a.update(elem.attrib)
        return a

    @staticmethod
    def _orientation(a, what):
        for bad in ("euler", "axisangle", "xyaxes", "zaxis", "fromto"):
            if bad in a:
                raise UnsupportedMJCF("%s orientation via %r" % (what, bad))
        q = _floats(a["quat"], 4) if "quat" in a else np.array([1.0, 0, 0, 0])
        return _normalize(q)

    # ---- bodies ----
    def parse_worldbody(self):
        world = dict(name="world", parent=-1, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                     geoms=[], joints=[], sites=[], children=[])
        for wb in self.root.findall("worldbody"):
            self._body_children(wb, world, None)
        self.world = world

    def _body_children(self, elem, body, childclass):
        for ch in elem:
            if ch.tag == "body":
                cc = ch.attrib.get("childclass", childclass)
                a = ch.attrib
                b = dict(name=a.get("name"), pos=_floats(a.get("pos", "0 0 0"), 3),
                         quat=self._orientation(a, "body"), geoms=[], joints=[], sites=[], children=[])
                body["children"].append(b)
                self._body_children(ch, b, cc)
            elif ch.tag == "geom":
                body["geoms"].append(self._geom(self._attrs(ch, "geom", childclass)))
            elif ch.tag == "joint":
                body["joints"].append(self._joint(self._attrs(ch, "joint", childclass)))
            elif ch.tag == "site":
                a = self._attrs(ch, "site", childclass)
                body["sites"].append(dict(name=a.get("name"), pos=_floats(a.get("pos", "0 0 0"), 3),
                                          quat=self._orientation(a, "site")))
            elif ch.tag == "composite":
                self._composite(ch, body)
            elif ch.tag in ("light", "camera"):
                pass
            elif ch.tag == "inertial":
                raise UnsupportedMJCF("<inertial> (the reference models use inertiafromgeom)")
            else:
                raise UnsupportedMJCF("<%s> inside a body" % ch.tag)

    def _geom(self, a, builtin_defaults=False):
        tname = a.get("type", "sphere")
        tmap = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE, "box": GEOM_BOX}
        if tname not in tmap:
            raise UnsupportedMJCF("geom type %r" % tname)
        size = np.zeros(3)
        sz = _floats(a.get("size", "0 0 0"))
        size[:sz.shape[0]] = sz
        fr = np.array([1.0, 0.005, 0.0001])
        if "friction" in a:
            f = _floats(a["friction"])
            fr[:f.shape[0]] = f
        return dict(name=a.get("name"), type=tmap[tname], size=size,
                    pos=_floats(a.get("pos", "0 0 0"), 3), quat=self._orientation(a, "geom"),
                    mass=float(a["mass"]) if "mass" in a else None,
                    density=float(a.get("density", 1000.0)),
                    contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
                    condim=int(a.get("condim", 3)), friction=fr,
                    solref=_floats(a.get("solref", "0.02 1"), 2),
                    solimp=self._solimp(a.get("solimp")),
                    margin=float(a.get("margin", 0)), gap=float(a.get("gap", 0)),
                    solmix=float(a.get("solmix", 1)))

    @staticmethod
    def _solimp(s):
        v = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
        if s is not None:
            f = _floats(s)
            v[:f.shape[0]] = f
        return v

    def _joint(self, a):
        tname = a.get("type", "hinge")
        if tname not in ("hinge", "slide"):
            raise UnsupportedMJCF("joint type %r (only hinge/slide are restated)" % tname)
        rng = _floats(a.get("range", "0 0"), 2)
        if tname == "hinge":
            rng = rng * self.angle_scale
        return dict(name=a.get("name"), type=JNT_HINGE if tname == "hinge" else JNT_SLIDE,
                    pos=_floats(a.get("pos", "0 0 0"), 3), axis=_normalize(_floats(a.get("axis", "0 0 1"), 3)),
                    limited=_bool(a.get("limited", "false")), range=rng,
                    stiffness=float(a.get("stiffness", 0)), damping=float(a.get("damping", 0)),
                    armature=float(a.get("armature", 0)), margin=float(a.get("margin", 0)),
                    ref=float(a.get("ref", 0)), springref=float(a.get("springref", 0)),
                    solref=_floats(a.get("solreflimit", "0.02 1"), 2),
                    solimp=self._solimp(a.get("solimplimit")),
                    frictionloss=float(a.get("frictionloss", 0)))

    # ---- composite (legacy box / ellipsoid shell; SURVEY App. A0, user_composite.cc MakeBox) ----
    def _composite(self, elem, body):
        a = elem.attrib
        prefix = a.get("prefix", "")
        ctype = a.get("type")
        if ctype not in ("box", "ellipsoid"):
            raise UnsupportedMJCF("composite type %r (only box/ellipsoid are restated)" % ctype)
        count = [int(x) for x in a["count"].split()]
        while len(count) < 3:
            count.append(1)
        if min(count) < 2:
            raise UnsupportedMJCF("composite box/ellipsoid needs count >= 2 in every dimension")
        spacing = float(a["spacing"])
        gattr, jattr, tattr = {}, None, None
        for ch in elem:
            if ch.tag == "geom":
                gattr = dict(ch.attrib)
            elif ch.tag == "joint":
                if ch.attrib.get("kind") != "main":
                    raise UnsupportedMJCF("composite joint kind %r" % ch.attrib.get("kind"))
                jattr = dict(ch.attrib)
            elif ch.tag == "tendon":
                if ch.attrib.get("kind") != "main":
                    raise UnsupportedMJCF("composite tendon kind %r" % ch.attrib.get("kind"))
                tattr = dict(ch.attrib)
            elif ch.tag == "skin":
                pass  # visual only
            else:
                raise UnsupportedMJCF("composite child <%s>" % ch.tag)
        jattr = jattr or {}
        tattr = tattr or {}
        # composite geoms start from MuJoCo's built-in defaults, not the model's default classes
        gel = self._geom(gattr)
        if gel["type"] not in (GEOM_CAPSULE, GEOM_SPHERE):
            gel["type"] = GEOM_SPHERE
        # centre geom: sphere, twice the element radius
        centre = dict(gel)
        centre.update(name=prefix + "Gcenter", type=GEOM_SPHERE,
                      size=np.array([2 * gel["size"][0], 0.0, 0.0]), pos=np.zeros(3),
                      quat=np.array([1.0, 0, 0, 0]))
        body["geoms"].append(centre)
        # the fixed "volume" tendon over all shell joints
        ten = dict(name=prefix + "T", type=TEN_FIXED, stiffness=float(tattr.get("stiffness", 0)),
                   damping=float(tattr.get("damping", 0)), wraps=[])
        self.tendons.append(ten)
        j_solref = _floats(jattr.get("solreffix", "0.02 1"), 2)
        j_solimp = self._solimp(jattr.get("solimpfix"))
        t_solref = _floats(tattr.get("solreffix", "0.02 1"), 2)
        t_solimp = self._solimp(tattr.get("solimpfix"))
        size = [0.5 * spacing * (c - 1) for c in count]

        def is_shell(ix, iy, iz):
            return (ix in (0, count[0] - 1)) or (iy in (0, count[1] - 1)) or (iz in (0, count[2] - 1))

        for ix in range(count[0]):
            for iy in range(count[1]):
                for iz in range(count[2]):
                    if not is_shell(ix, iy, iz):
                        continue
                    pos = np.array([2.0 * ix / (count[0] - 1) - 1, 2.0 * iy / (count[1] - 1) - 1,
                                    2.0 * iz / (count[2] - 1) - 1])
                    if ctype == "ellipsoid":
                        pos = _normalize(pos)
                    pos = pos * size
                    suffix = "%d_%d_%d" % (ix, iy, iz)
                    g = dict(gel)
                    g.update(name=prefix + "G" + suffix, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]))
                    if g["type"] == GEOM_CAPSULE:
                        g["pos"] = np.array([0, 0, -(g["size"][0] + g["size"][1])])
                    else:
                        g["pos"] = np.array([0, 0, -g["size"][0]])
                    jname = prefix + "J" + suffix
                    jn = self._joint({"type": "slide", "axis": "0 0 1", "name": jname,
                                      **{k: v for k, v in jattr.items()
                                         if k in ("stiffness", "damping", "armature", "limited",
                                                  "range", "margin", "frictionloss")}})
                    b = dict(name=prefix + "B" + suffix, pos=pos, quat=z2quat(pos), geoms=[g],
                             joints=[jn], sites=[], children=[])
                    body["children"].append(b)
                    self.equalities.append(dict(type=EQ_JOINT, name1=jname, name2=None,
                                                solref=j_solref, solimp=j_solimp))
                    ten["wraps"].append((jname, 1.0))
                    for k in range(3):
                        n = [min(ix + (k == 0), count[0] - 1), min(iy + (k == 1), count[1] - 1),
                             min(iz + (k == 2), count[2] - 1)]
                        if is_shell(*n) and tuple(n) != (ix, iy, iz):
                            self.equalities.append(dict(type=EQ_JOINT, name1=jname,
                                                        name2=prefix + "J%d_%d_%d" % tuple(n),
                                                        solref=j_solref, solimp=j_solimp))
        self.equalities.append(dict(type=EQ_TENDON, name1=ten["name"], name2=None,
                                    solref=t_solref, solimp=t_solimp))

    # ---- tendon / actuator / sensor sections ----
    def parse_sections(self):
        if self.root.find("equality") is not None or self.root.find("contact") is not None:
            raise UnsupportedMJCF("<equality>/<contact> sections are not used by the reference models")
        for sec in self.root.findall("tendon"):
            for t in sec:
                if t.tag != "spatial":
                    raise UnsupportedMJCF("tendon <%s>" % t.tag)
                a = dict(self.defaults["main"]["tendon"])
                a.update(t.attrib)
                wraps = []
                for w in t:
                    if w.tag != "site":
                        raise UnsupportedMJCF("spatial tendon wrap <%s>" % w.tag)
                    wraps.append(w.attrib["site"])
                if len(wraps) != 2:
                    raise UnsupportedMJCF("spatial tendons with exactly two sites only")
                self.tendons.append(dict(name=a.get("name"), type=TEN_SPATIAL,
                                         stiffness=float(a.get("stiffness", 0)),
                                         damping=float(a.get("damping", 0)), wraps=wraps))
        for sec in self.root.findall("actuator"):
            for act in sec:
                if act.tag != "cylinder":
                    raise UnsupportedMJCF("actuator <%s>" % act.tag)
                a = dict(self.defaults["main"]["cylinder"])
                a.update(act.attrib)
                if "tendon" not in a:
                    raise UnsupportedMJCF("cylinder actuators must act on a tendon")
                if _bool(a.get("ctrllimited", "false")) or _bool(a.get("forcelimited", "false")):
                    raise UnsupportedMJCF("actuator limits")
                area = float(a["area"]) if "area" in a else math.pi / 4 * float(a.get("diameter", 0)) ** 2
                bias = np.zeros(3)
                if "bias" in a:
                    bias = _floats(a["bias"], 3)
                self.actuators.append(dict(tendon=a["tendon"], gear=float(a.get("gear", "1").split()[0]),
                                           timeconst=float(a.get("timeconst", 1)), area=area, bias=bias))
        for sec in self.root.findall("sensor"):
            for s in sec:
                if s.tag not in ("accelerometer", "gyro"):
                    raise UnsupportedMJCF("sensor <%s>" % s.tag)
                if float(s.attrib.get("noise", 0)) != 0 or float(s.attrib.get("cutoff", 0)) != 0:
                    raise UnsupportedMJCF("sensor noise/cutoff")
                self.sensors.append(dict(type=SENS_ACCEL if s.tag == "accelerometer" else SENS_GYRO,
                                         site=s.attrib["site"], name=s.attrib.get("name")))

    # ---- flatten ----
    def build(self, name):
        self.parse_worldbody()
        self.parse_sections()
        A = {}
        names = {"body": [], "geom": [], "joint": [], "site": [], "tendon": []}
        body_parent, body_pos, body_quat = [], [], []
        body_jntadr, body_jntnum, body_geomadr, body_geomnum = [], [], [], []
        geoms, joints, sites = [], [], []

        def rec(b, parent):
            bid = len(body_parent)
            body_parent.append(parent)
            body_pos.append(b["pos"])
            body_quat.append(b["quat"])
            names["body"].append(b["name"])
            body_jntadr.append(len(joints))
            body_jntnum.append(len(b["joints"]))
            body_geomadr.append(len(geoms))
            body_geomnum.append(len(b["geoms"]))
            for j in b["joints"]:
                joints.append((bid, j))
            for g in b["geoms"]:
                geoms.append((bid, g))
            for s in b["sites"]:
                sites.append((bid, s))
            for c in b["children"]:
                rec(c, bid)

        rec(self.world, -1)
        body_parent[0] = 0
        nbody, njnt, ngeom, nsite = len(body_parent), len(joints), len(geoms), len(sites)
        nv = njnt
        if joints and joints[0][0] == 0:
            raise UnsupportedMJCF("joints in the world body")

        A["body_parentid"] = np.array(body_parent, dtype=np.int32)
        A["body_pos"] = np.array(body_pos, dtype=np.float64).reshape(nbody, 3)
        A["body_quat"] = np.array(body_quat, dtype=np.float64).reshape(nbody, 4)
        A["body_jntadr"] = np.array(body_jntadr, dtype=np.int32)
        A["body_jntnum"] = np.array(body_jntnum, dtype=np.int32)
        A["body_dofadr"] = A["body_jntadr"].copy()      # 1 dof per joint in this subset
        A["body_dofnum"] = A["body_jntnum"].copy()
        A["body_geomadr"] = np.array(body_geomadr, dtype=np.int32)
        A["body_geomnum"] = np.array(body_geomnum, dtype=np.int32)
        # weld id / root id
        weld = np.zeros(nbody, dtype=np.int32)
        rootid = np.zeros(nbody, dtype=np.int32)
        for i in range(1, nbody):
            p = body_parent[i]
            weld[i] = i if body_jntnum[i] > 0 else weld[p]
            rootid[i] = i if p == 0 else rootid[p]
        A["body_weldid"] = weld
        A["body_rootid"] = rootid

        # joints / dofs
        names["joint"] = [j["name"] for _, j in joints]
        A["jnt_type"] = np.array([j["type"] for _, j in joints], dtype=np.int32)
        A["jnt_bodyid"] = np.array([b for b, _ in joints], dtype=np.int32)
        A["jnt_pos"] = np.array([j["pos"] for _, j in joints]).reshape(njnt, 3)
        A["jnt_axis"] = np.array([j["axis"] for _, j in joints]).reshape(njnt, 3)
        A["jnt_limited"] = np.array([int(j["limited"]) for _, j in joints], dtype=np.int32)
        A["jnt_range"] = np.array([j["range"] for _, j in joints]).reshape(njnt, 2)
        A["jnt_stiffness"] = np.array([j["stiffness"] for _, j in joints], dtype=np.float64)
        A["jnt_margin"] = np.array([j["margin"] for _, j in joints], dtype=np.float64)
        A["jnt_solref"] = np.array([j["solref"] for _, j in joints]).reshape(njnt, 2)
        A["jnt_solimp"] = np.array([j["solimp"] for _, j in joints]).reshape(njnt, 5)
        A["qpos0"] = np.array([j["ref"] for _, j in joints], dtype=np.float64)
        A["qpos_spring"] = np.array([j["springref"] for _, j in joints], dtype=np.float64)
        for _, j in joints:
            if j["frictionloss"] != 0 or j["armature"] != 0:
                raise UnsupportedMJCF("joint frictionloss/armature")
        A["dof_bodyid"] = A["jnt_bodyid"].copy()
        A["dof_damping"] = np.array([j["damping"] for _, j in joints], dtype=np.float64)
        dof_parent = np.full(nv, -1, dtype=np.int32)
        last_dof_of_body = {}
        for d in range(nv):
            b = int(A["dof_bodyid"][d])
            if d > 0 and A["dof_bodyid"][d - 1] == b:
                dof_parent[d] = d - 1
            else:
                p = body_parent[b]
                while p > 0 and body_jntnum[p] == 0:
                    p = body_parent[p]
                dof_parent[d] = last_dof_of_body.get(p, -1) if p > 0 else -1
            last_dof_of_body[b] = d
        A["dof_parentid"] = dof_parent
        madr = np.zeros(nv, dtype=np.int32)
        nM = 0
        for d in range(nv):
            madr[d] = nM
            k = d
            while k >= 0:
                nM += 1
                k = dof_parent[k]
        A["dof_Madr"] = madr
        self.nM = nM

        # geoms
        names["geom"] = [g["name"] for _, g in geoms]
        A["geom_type"] = np.array([g["type"] for _, g in geoms], dtype=np.int32)
        A["geom_bodyid"] = np.array([b for b, _ in geoms], dtype=np.int32)
        A["geom_pos"] = np.array([g["pos"] for _, g in geoms]).reshape(ngeom, 3)
        A["geom_quat"] = np.array([g["quat"] for _, g in geoms]).reshape(ngeom, 4)
        A["geom_size"] = np.array([g["size"] for _, g in geoms]).reshape(ngeom, 3)
        A["geom_contype"] = np.array([g["contype"] for _, g in geoms], dtype=np.int32)
        A["geom_conaffinity"] = np.array([g["conaffinity"] for _, g in geoms], dtype=np.int32)
        A["geom_condim"] = np.array([g["condim"] for _, g in geoms], dtype=np.int32)
        A["geom_friction"] = np.array([g["friction"] for _, g in geoms]).reshape(ngeom, 3)
        A["geom_solref"] = np.array([g["solref"] for _, g in geoms]).reshape(ngeom, 2)
        A["geom_solimp"] = np.array([g["solimp"] for _, g in geoms]).reshape(ngeom, 5)
        A["geom_margin"] = np.array([g["margin"] for _, g in geoms], dtype=np.float64)
        A["geom_gap"] = np.array([g["gap"] for _, g in geoms], dtype=np.float64)
        A["geom_solmix"] = np.array([g["solmix"] for _, g in geoms], dtype=np.float64)
        rb = np.zeros(ngeom)
        for i, (_, g) in enumerate(geoms):
            s = g["size"]
            rb[i] = {GEOM_PLANE: 0.0, GEOM_SPHERE: s[0], GEOM_CAPSULE: s[0] + s[1],
                     GEOM_BOX: float(np.linalg.norm(s))}[g["type"]]
        A["geom_rbound"] = rb

        # sites
        names["site"] = [s["name"] for _, s in sites]
        A["site_bodyid"] = np.array([b for b, _ in sites], dtype=np.int32)
        A["site_pos"] = np.array([s["pos"] for _, s in sites]).reshape(nsite, 3)
        A["site_quat"] = np.array([s["quat"] for _, s in sites]).reshape(nsite, 4)

        # mass properties from geoms
        bmass = np.zeros(nbody)
        binertia = np.zeros((nbody, 3))
        bipos = np.zeros((nbody, 3))
        biquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
        for b in range(nbody):
            gl = [g for bb, g in geoms if bb == b]
            props = [self._geom_mass(g) for g in gl]
            props = [(g, m, I) for g, (m, I) in zip(gl, props) if m > 0]
            if not props:
                continue
            if len(props) == 1:
                g, m, I = props[0]
                bmass[b], binertia[b], bipos[b], biquat[b] = m, I, g["pos"], g["quat"]
            else:
                mt = sum(m for _, m, _ in props)
                com = sum(m * g["pos"] for g, m, _ in props) / mt
                It = np.zeros((3, 3))
                for g, m, I in props:
                    R = quat_to_mat(g["quat"])
                    d = g["pos"] - com
                    It += R @ np.diag(I) @ R.T + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
                w, V = np.linalg.eigh(It)
                order = np.argsort(-w)
                w, V = w[order], V[:, order]
                if np.linalg.det(V) < 0:
                    V[:, 2] = -V[:, 2]
                bmass[b], binertia[b], bipos[b] = mt, w, com
                biquat[b] = _mat_to_quat(V)
        if self.settotalmass > 0:
            scale = self.settotalmass / bmass[1:].sum()
            bmass *= scale
            binertia *= scale
            self.mass_scale = scale
        else:
            self.mass_scale = 1.0
        A["body_mass"], A["body_inertia"], A["body_ipos"], A["body_iquat"] = bmass, binertia, bipos, biquat
        for b in range(1, nbody):
            if body_jntnum[b] > 0 and bmass[b] <= 0:
                raise UnsupportedMJCF("moving body %d has no mass" % b)

        # tendons (composite tendon first, then the <tendon> section)
        jid = {n: i for i, n in enumerate(names["joint"]) if n is not None}
        sid = {n: i for i, n in enumerate(names["site"]) if n is not None}
        nten = len(self.tendons)
        names["tendon"] = [t["name"] for t in self.tendons]
        A["tendon_type"] = np.array([t["type"] for t in self.tendons], dtype=np.int32)
        A["tendon_stiffness"] = np.array([t["stiffness"] for t in self.tendons], dtype=np.float64)
        A["tendon_damping"] = np.array([t["damping"] for t in self.tendons], dtype=np.float64)
        wadr, wnum, wobj, wprm = [], [], [], []
        for t in self.tendons:
            wadr.append(len(wobj))
            wnum.append(len(t["wraps"]))
            for w in t["wraps"]:
                if t["type"] == TEN_FIXED:
                    wobj.append(jid[w[0]])
                    wprm.append(w[1])
                else:
                    wobj.append(sid[w])
                    wprm.append(0.0)
        A["tendon_adr"] = np.array(wadr, dtype=np.int32)
        A["tendon_num"] = np.array(wnum, dtype=np.int32)
        A["wrap_objid"] = np.array(wobj, dtype=np.int32)
        A["wrap_prm"] = np.array(wprm, dtype=np.float64)

        # equalities
        tid = {n: i for i, n in enumerate(names["tendon"]) if n is not None}
        neq = len(self.equalities)
        A["eq_type"] = np.array([e["type"] for e in self.equalities], dtype=np.int32)
        o1, o2 = [], []
        for e in self.equalities:
            table = jid if e["type"] == EQ_JOINT else tid
            o1.append(table[e["name1"]])
            o2.append(table[e["name2"]] if e["name2"] is not None else -1)
        A["eq_obj1id"] = np.array(o1, dtype=np.int32)
        A["eq_obj2id"] = np.array(o2, dtype=np.int32)
        data = np.zeros((neq, 5))
        data[:, 1] = 1.0                                   # polycoef "0 1 0 0 0"
        A["eq_data"] = data
        A["eq_solref"] = np.array([e["solref"] for e in self.equalities]).reshape(neq, 2)
        A["eq_solimp"] = np.array([e["solimp"] for e in self.equalities]).reshape(neq, 5)

        # actuators (cylinder: filter dynamics, fixed gain = area, affine bias)
        nu = len(self.actuators)
        A["actuator_trnid"] = np.array([tid[a["tendon"]] for a in self.actuators], dtype=np.int32)
        A["actuator_gear"] = np.array([a["gear"] for a in self.actuators], dtype=np.float64)
        A["actuator_timeconst"] = np.array([a["timeconst"] for a in self.actuators], dtype=np.float64)
        A["actuator_gain"] = np.array([a["area"] for a in self.actuators], dtype=np.float64)
        A["actuator_bias"] = np.array([a["bias"] for a in self.actuators]).reshape(nu, 3)

        # sensors
        ns = len(self.sensors)
        A["sensor_type"] = np.array([s["type"] for s in self.sensors], dtype=np.int32)
        A["sensor_objid"] = np.array([sid[s["site"]] for s in self.sensors], dtype=np.int32)
        A["sensor_adr"] = np.arange(ns, dtype=np.int32) * 3

        m = Model(name=name, opt=dict(self.opt), arrays=A, names=names)
        m.opt["nM"] = nM
        m.opt["mass_scale"] = self.mass_scale
        _set_const(m)
        return m

    @staticmethod
    def _geom_mass(g):
        s = g["size"]
        t = g["type"]
        if t == GEOM_PLANE:
            return 0.0, np.zeros(3)
        if t == GEOM_BOX:
            vol = 8 * s[0] * s[1] * s[2]
        elif t == GEOM_SPHERE:
            vol = 4.0 / 3.0 * math.pi * s[0] ** 3
        else:
            vol = math.pi * s[0] ** 2 * (2 * s[1]) + 4.0 / 3.0 * math.pi * s[0] ** 3
        mass = g["mass"] if g["mass"] is not None else g["density"] * vol
        if t == GEOM_BOX:
            I = mass / 3.0 * np.array([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2])
        elif t == GEOM_SPHERE:
            I = np.full(3, 0.4 * mass * s[0] ** 2)
        else:
            r, h = s[0], 2 * s[1]
            ms = mass * (4.0 / 3.0 * math.pi * r ** 3) / vol
            mc = mass - ms
            Ix = mc * (3 * r * r + h * h) / 12.0 + 0.4 * ms * r * r + ms * h * (3 * r + 2 * h) / 8.0
            I = np.array([Ix, Ix, mc * r * r / 2.0 + 0.4 * ms * r * r])
        return mass, I


def _mat_to_quat(R):
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return _normalize(q)


# --------------------------------------------------------------------------------------------
# numpy forward kinematics + dense inertia (used for the compile-time constants only)
# --------------------------------------------------------------------------------------------
def kinematics(m: Model, qpos):
    """World frames of bodies, joints, geoms and sites at ``qpos``."""
    A = m.arrays
    nb = m.nbody
    xpos = np.zeros((nb, 3))
    xquat = np.tile(np.array([1.0, 0, 0, 0]), (nb, 1))
    nj = A["jnt_type"].shape[0]
    xanchor = np.zeros((nj, 3))
    xaxis = np.zeros((nj, 3))
    for b in range(1, nb):
        p = A["body_parentid"][b]
        Rp = quat_to_mat(xquat[p])
        pos = xpos[p] + Rp @ A["body_pos"][b]
        quat = quat_mul(xquat[p], A["body_quat"][b])
        for j in range(A["body_jntadr"][b], A["body_jntadr"][b] + A["body_jntnum"][b]):
            R = quat_to_mat(quat)
            xaxis[j] = R @ A["jnt_axis"][j]
            xanchor[j] = pos + R @ A["jnt_pos"][j]
            dq = qpos[j] - A["qpos0"][j]
            if A["jnt_type"][j] == JNT_SLIDE:
                pos = pos + xaxis[j] * dq
            else:
                quat = quat_mul(quat, axis_angle_quat(A["jnt_axis"][j], dq))
                pos = xanchor[j] - quat_to_mat(quat) @ A["jnt_pos"][j]
        xpos[b] = pos
        xquat[b] = _normalize(quat)
    xmat = np.array([quat_to_mat(q) for q in xquat])
    out = dict(xpos=xpos, xquat=xquat, xmat=xmat, xanchor=xanchor, xaxis=xaxis)
    out["xipos"] = xpos + np.einsum("bij,bj->bi", xmat, A["body_ipos"])
    out["ximat"] = np.array([xmat[b] @ quat_to_mat(A["body_iquat"][b]) for b in range(nb)])
    gb = A["geom_bodyid"]
    out["geom_xpos"] = xpos[gb] + np.einsum("gij,gj->gi", xmat[gb], A["geom_pos"])
    out["geom_xmat"] = np.array([xmat[gb[g]] @ quat_to_mat(A["geom_quat"][g]) for g in range(m.ngeom)])
    sb = A["site_bodyid"]
    out["site_xpos"] = xpos[sb] + np.einsum("sij,sj->si", xmat[sb], A["site_pos"])
    out["site_xmat"] = np.array([xmat[sb[s]] @ quat_to_mat(A["site_quat"][s]) for s in range(sb.shape[0])])
    return out


def jacobian(m: Model, kin, body, point):
    """3 x nv translational and rotational Jacobians of ``point`` attached to ``body``."""
    A = m.arrays
    nv = m.nv
    jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
    b = body
    while b > 0:
        for j in range(A["body_jntadr"][b], A["body_jntadr"][b] + A["body_jntnum"][b]):
            if A["jnt_type"][j] == JNT_SLIDE:
                jp[:, j] = kin["xaxis"][j]
            else:
                jr[:, j] = kin["xaxis"][j]
                jp[:, j] = np.cross(kin["xaxis"][j], point - kin["xanchor"][j])
        b = A["body_parentid"][b]
    return jp, jr


def dense_inertia(m: Model, kin):
    A = m.arrays
    nv = m.nv
    M = np.zeros((nv, nv))
    for b in range(1, m.nbody):
        if A["body_mass"][b] <= 0 or A["body_weldid"][b] == 0:
            continue
        jp, jr = jacobian(m, kin, b, kin["xipos"][b])
        Iw = kin["ximat"][b] @ np.diag(A["body_inertia"][b]) @ kin["ximat"][b].T
        M += A["body_mass"][b] * jp.T @ jp + jr.T @ Iw @ jr
    return M


def tendon_length_jac(m: Model, kin, qpos):
    A = m.arrays
    nt, nv = m.ntendon, m.nv
    L = np.zeros(nt)
    J = np.zeros((nt, nv))
    for t in range(nt):
        adr, num = A["tendon_adr"][t], A["tendon_num"][t]
        if A["tendon_type"][t] == TEN_FIXED:
            for w in range(adr, adr + num):
                j = A["wrap_objid"][w]
                L[t] += A["wrap_prm"][w] * qpos[j]
                J[t, j] = A["wrap_prm"][w]
        else:
            s0, s1 = A["wrap_objid"][adr], A["wrap_objid"][adr + 1]
            p0, p1 = kin["site_xpos"][s0], kin["site_xpos"][s1]
            d = p1 - p0
            L[t] = np.linalg.norm(d)
            jp0, _ = jacobian(m, kin, A["site_bodyid"][s0], p0)
            jp1, _ = jacobian(m, kin, A["site_bodyid"][s1], p1)
            J[t] = (d / L[t]) @ (jp1 - jp0)
    return L, J


def _set_const(m: Model):
    """Compile-time constants of ``mj_setConst`` at qpos0 (SURVEY App. A0 last bullet)."""
    A = m.arrays
    nv = m.nv
    kin = kinematics(m, A["qpos0"])
    M = dense_inertia(m, kin)
    Minv = np.linalg.inv(M)
    A["dof_invweight0"] = np.diag(Minv).copy()
    biw = np.zeros((m.nbody, 2))
    for b in range(1, m.nbody):
        if A["body_weldid"][b] == 0:
            continue
        jp, jr = jacobian(m, kin, b, kin["xipos"][b])
        biw[b, 0] = np.trace(jp @ Minv @ jp.T) / 3.0
        biw[b, 1] = np.trace(jr @ Minv @ jr.T) / 3.0
    A["body_invweight0"] = biw
    L, J = tendon_length_jac(m, kin, A["qpos0"])
    A["tendon_length0"] = L
    A["tendon_lengthspring"] = L.copy()
    A["tendon_invweight0"] = np.einsum("ti,ij,tj->t", J, Minv, J)
    m.opt["meaninertia"] = float(np.trace(M) / max(1, nv))
    # subtree masses (used by comPos)
    sub = A["body_mass"].copy()
    for b in range(m.nbody - 1, 0, -1):
        sub[A["body_parentid"][b]] += sub[b]
    A["body_subtreemass"] = sub


# --------------------------------------------------------------------------------------------
# public entry points
# --------------------------------------------------------------------------------------------
def load_mjcf(path) -> Model:
    """Compile an MJCF file of the supported subset (the replacement for
    ``mujoco_py.load_model_from_path``, ref: environment/manenv.py:27)."""
    root = _load_spliced(path)
    return _Builder(root).build(root.attrib.get("model", os.path.basename(path)))


# blob format: "SGM1" | u32 version | u32 nsections | sections[name[32] | u32 dtype | u32 count | u64 offset]
# dtype: 0 = float64, 1 = int32.  Data follows the table, each section 8-byte aligned.
_MAGIC = b"SGM1"
_VERSION = 1
_OPT_KEYS = ("timestep", "gx", "gy", "gz", "iterations", "tolerance", "impratio", "meaninertia",
             "nconmax", "njmax", "nM", "mass_scale")


def model_to_blob(m: Model) -> bytes:
    opt = m.opt
    optv = np.array([opt["timestep"], opt["gravity"][0], opt["gravity"][1], opt["gravity"][2],
                     opt["iterations"], opt["tolerance"], opt["impratio"], opt["meaninertia"],
                     opt["nconmax"], opt["njmax"], opt["nM"], opt["mass_scale"]], dtype=np.float64)
    secs = [("opt", optv)]
    for k in sorted(m.arrays):
        a = m.arrays[k]
        if a.dtype.kind == "i":
            a = np.ascontiguousarray(a, dtype=np.int32)
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
        secs.append((k, a))
    # names of geoms travel too (contact-name scan of get_sensor_sensordata, ref: manenv.py:71-80)
    gn = "\0".join((n or "") for n in m.names["geom"]).encode() + b"\0"
    secs.append(("geom_names", np.frombuffer(gn + b"\0" * ((-len(gn)) % 4), dtype=np.int32)))
    header = struct.calcsize("<4sII")
    entry = struct.calcsize("<32sIIQ")
    off = header + entry * len(secs)
    off += (-off) % 8
    table, payload = b"", b""
    for name, a in secs:
        raw = a.tobytes()
        table += struct.pack("<32sIIQ", name.encode(), 0 if a.dtype == np.float64 else 1, a.size, off + len(payload))
        payload += raw + b"\0" * ((-len(raw)) % 8)
    head = struct.pack("<4sII", _MAGIC, _VERSION, len(secs)) + table
    head += b"\0" * ((-len(head)) % 8)
    return head + payload


def save_blob(m: Model, path):
    with open(path, "wb") as f:
        f.write(model_to_blob(m))


_SHAPES = {"body_pos": 3, "body_quat": 4, "body_ipos": 3, "body_iquat": 4, "body_inertia": 3,
           "body_invweight0": 2, "jnt_pos": 3, "jnt_axis": 3, "jnt_range": 2, "jnt_solref": 2,
           "jnt_solimp": 5, "geom_pos": 3, "geom_quat": 4, "geom_size": 3, "geom_friction": 3,
           "geom_solref": 2, "geom_solimp": 5, "site_pos": 3, "site_quat": 4, "eq_data": 5,
           "eq_solref": 2, "eq_solimp": 5, "actuator_bias": 3}


def load_blob(data) -> Model:
    """Inverse of :func:`model_to_blob` (path or bytes)."""
    if isinstance(data, (str, os.PathLike)):
        with open(data, "rb") as f:
            data = f.read()
    magic, ver, n = struct.unpack_from("<4sII", data, 0)
    if magic != _MAGIC or ver != _VERSION:
        raise ValueError("not a softgrip model blob")
    off = struct.calcsize("<4sII")
    arrays, opt, names = {}, {}, {"geom": []}
    for _ in range(n):
        name, dt, cnt, o = struct.unpack_from("<32sIIQ", data, off)
        off += struct.calcsize("<32sIIQ")
        name = name.rstrip(b"\0").decode()
        a = np.frombuffer(data, dtype=np.float64 if dt == 0 else np.int32, count=cnt, offset=o).copy()
        if name == "opt":
            v = dict(zip(_OPT_KEYS, a))
            opt = {"timestep": v["timestep"], "gravity": np.array([v["gx"], v["gy"], v["gz"]]),
                   "iterations": int(v["iterations"]), "tolerance": v["tolerance"],
                   "impratio": v["impratio"], "meaninertia": v["meaninertia"],
                   "nconmax": int(v["nconmax"]), "njmax": int(v["njmax"]), "nM": int(v["nM"]),
                   "mass_scale": v["mass_scale"]}
        elif name == "geom_names":
            names["geom"] = [s or None for s in a.tobytes().rstrip(b"\0").decode().split("\0")]
        else:
            if name in _SHAPES:
                a = a.reshape(-1, _SHAPES[name])
            arrays[name] = a
    m = Model(name="blob", opt=opt, arrays=arrays, names=names)
    ng = m.ngeom
    names["geom"] = (names["geom"] + [None] * ng)[:ng]
    return m
