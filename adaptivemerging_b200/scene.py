"""Scene hand-over: the structure-of-arrays blob the C ABI ingests (``am3d_scene``).

In a deployment the reference's own Java ``XMLParser`` (kept intact)
fills this blob from its ``RigidBody`` objects.  There is no JVM in this repo's test
environment, so this module re-reads the same scene XML / ``.sph`` files and reproduces the
loader arithmetic of ``src/mergingBodies3D/XMLParser.java`` (createBox :360-396, createPlane
:404-413, createSphere :420-447, createMesh :449-539, createComposite :180-353,
setCommonAttributes :546-634) together with the ``javax.vecmath`` 1.3.2 routines it calls
(``Matrix3d.set(AxisAngle4d)`` jar Matrix3d.java:868-911, ``invert`` :1067-1133 LU with
implicit-scaling partial pivoting).  It also holds the synthetic scene generators for the
BASELINE.json configurations (1M-box stack / pile, 100k sphere-tree funnel, batched copies).

Nothing here runs on the per-step path; it is the caller side of the boundary.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

from .volint import mesh_mass_properties  # VolInt / Polyhedron / PolygonSoup arithmetic (XMLParser.createMesh :457-464)

BODY_BOX, BODY_PLANE, BODY_SPHERE, BODY_MESH, BODY_COMPOSITE = 0, 1, 2, 3, 4
SHAPE_BOX, SHAPE_TREE, SHAPE_PLANE = 0, 1, 2
F_PINNED, F_MAGNETIC, F_MAGNET_ACTIVE, F_SLEEPING = 1, 2, 4, 8
SPRING_ZERO, SPRING_WORLD, SPRING_BODYBODY = 0, 1, 2

_DMAX = 1.7976931348623157e308
_DMIN = 5e-324  # Java Double.MIN_VALUE, used as "-inf" by the reference (XMLParser.java:281, 476)


# ---------------------------------------------------------------------------------------------
# javax.vecmath 1.3.2 restatements (scalar, operation order preserved)
# ---------------------------------------------------------------------------------------------
def mat_from_axis_angle(ax, ay, az, angle):
    """Matrix3d.set(AxisAngle4d): off-diagonals use the UN-normalised axis products."""
    mag = math.sqrt(ax * ax + ay * ay + az * az)
    if mag < 1.110223024e-16:
        return np.eye(3)
    mag = 1.0 / mag
    nx, ny, nz = ax * mag, ay * mag, az * mag
    s, c = math.sin(angle), math.cos(angle)
    t = 1.0 - c
    xz, xy, yz = ax * az, ax * ay, ay * az
    m = np.empty((3, 3))
    m[0, 0] = t * nx * nx + c
    m[0, 1] = t * xy - s * nz
    m[0, 2] = t * xz + s * ny
    m[1, 0] = t * xy + s * nz
    m[1, 1] = t * ny * ny + c
    m[1, 2] = t * yz - s * nx
    m[2, 0] = t * xz - s * ny
    m[2, 1] = t * yz + s * nx
    m[2, 2] = t * nz * nz + c
    return m


def mat_mul(a, b):
    """Matrix3d.mul(m1,m2): plain left-to-right three-term sums."""
    r = np.empty((3, 3))
    for i in range(3):
        for j in range(3):
            r[i, j] = a[i, 0] * b[0, j] + a[i, 1] * b[1, j] + a[i, 2] * b[2, j]
    return r


def mat_mul_transpose_right(a, b):
    r = np.empty((3, 3))
    for i in range(3):
        for j in range(3):
            r[i, j] = a[i, 0] * b[j, 0] + a[i, 1] * b[j, 1] + a[i, 2] * b[j, 2]
    return r


def rm0rt(R, M):
    """RigidTransform3D.computeRM0RT (RigidTransform3D.java:229-232)."""
    return mat_mul_transpose_right(mat_mul(R, M), R)


def mat_transform(m, v):
    return np.array([m[0, 0] * v[0] + m[0, 1] * v[1] + m[0, 2] * v[2],
                     m[1, 0] * v[0] + m[1, 1] * v[1] + m[1, 2] * v[2],
                     m[2, 0] * v[0] + m[2, 1] * v[1] + m[2, 2] * v[2]])


def lu_invert3(m):
    """Matrix3d.invertGeneral: Crout LU with implicit row scaling + back substitution."""
    a = [float(x) for x in np.asarray(m, dtype=np.float64).reshape(9)]
    row_scale = [0.0] * 3
    for i in range(3):
        big = max(abs(a[3 * i]), abs(a[3 * i + 1]), abs(a[3 * i + 2]))
        if big == 0.0:
            raise ZeroDivisionError("singular matrix")
        row_scale[i] = 1.0 / big
    perm = [0, 0, 0]
    for j in range(3):
        for i in range(j):
            s = a[3 * i + j]
            for k in range(i):
                s -= a[3 * i + k] * a[3 * k + j]
            a[3 * i + j] = s
        big = 0.0
        imax = -1
        for i in range(j, 3):
            s = a[3 * i + j]
            for k in range(j):
                s -= a[3 * i + k] * a[3 * k + j]
            a[3 * i + j] = s
            t = row_scale[i] * abs(s)
            if t >= big:
                big = t
                imax = i
        if j != imax:
            for k in range(3):
                a[3 * imax + k], a[3 * j + k] = a[3 * j + k], a[3 * imax + k]
            row_scale[imax] = row_scale[j]
        perm[j] = imax
        if a[3 * j + j] == 0.0:
            raise ZeroDivisionError("singular matrix")
        if j != 2:
            t = 1.0 / a[3 * j + j]
            for i in range(j + 1, 3):
                a[3 * i + j] *= t
    r = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]
    for k in range(3):
        ii = -1
        for i in range(3):
            ip = perm[i]
            s = r[k + 3 * ip]
            r[k + 3 * ip] = r[k + 3 * i]
            if ii >= 0:
                for j in range(ii, i):
                    s -= a[3 * i + j] * r[k + 3 * j]
            elif s != 0.0:
                ii = i
            r[k + 3 * i] = s
        r[k + 6] /= a[8]
        r[k + 3] = (r[k + 3] - a[5] * r[k + 6]) / a[4]
        r[k] = (r[k] - a[1] * r[k + 3] - a[2] * r[k + 6]) / a[0]
    return np.array(r).reshape(3, 3)


# ---------------------------------------------------------------------------------------------
# loader-side objects
# ---------------------------------------------------------------------------------------------
@dataclass
class Tree:
    """A sphere tree in breadth-first layout (children contiguous)."""
    c: np.ndarray            # [n,3]
    r: np.ndarray            # [n]
    first_child: np.ndarray  # [n] local index or -1
    child_count: np.ndarray  # [n]
    rank: np.ndarray         # [n] pre-order DFS index


@dataclass
class Part:
    """One collision primitive."""
    type: int
    size: np.ndarray = field(default_factory=lambda: np.zeros(3))
    radius: float = 0.0
    p: np.ndarray = field(default_factory=lambda: np.zeros(3))
    B2C_R: np.ndarray = field(default_factory=lambda: np.eye(3))
    B2C_t: np.ndarray = field(default_factory=lambda: np.zeros(3))
    tree: int = -1  # index into SceneBuilder.trees


@dataclass
class Body:
    name: str
    type: int
    mass: float = 0.0
    minv: float = 0.0
    mass_angular0: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    jinv0: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    x: np.ndarray = field(default_factory=lambda: np.zeros(3))
    R: np.ndarray = field(default_factory=lambda: np.eye(3))
    v: np.ndarray = field(default_factory=lambda: np.zeros(3))
    omega: np.ndarray = field(default_factory=lambda: np.zeros(3))
    pinned: bool = False
    magnetic: bool = False
    friction: float = 0.8
    restitution: float = 0.0
    bbB: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    parts: list = field(default_factory=list)
    scene: int = 0


@dataclass
class SpringDef:
    type: int
    body1: int
    body2: int
    pb1: np.ndarray
    pb2: np.ndarray
    pw: np.ndarray
    k: float = 100.0
    d: float = 10.0
    l0: float = 0.5
    ls: float = 1.0


def _floats(s):
    return [float(t) for t in s.strip().split()]


def _new_rigid_body(name, btype, mass, mass_angular, pinned, bbB):
    """RigidBody(massLinear, massAngular, pinned, boundingBoxB) (RigidBody.java:182-206)."""
    b = Body(name=name, type=btype)
    b.pinned = pinned
    if not pinned:
        b.bbB = np.array(bbB, dtype=np.float64).reshape(-1, 3)
        b.mass = mass
        b.minv = 1.0 / mass
        if mass_angular is not None:
            b.mass_angular0 = np.array(mass_angular, dtype=np.float64)
            b.jinv0 = lu_invert3(b.mass_angular0)
    return b


def read_sph(path, scale, com):
    """The .sph reader of XMLParser.createMesh (:500-535): scale, then subtract the COM.
    Returns a breadth-first Tree (the reference links nodes by 1-based index)."""
    with open(path) as f:
        lines = [ln for ln in f.read().splitlines()]
    n = int(lines[0].split()[0])
    cB = np.empty((n, 3))
    rr = np.empty(n)
    for i in range(n):
        t = lines[1 + i].split()
        p = np.array([float(t[0]), float(t[1]), float(t[2])])
        p = p * scale
        p = p - com
        cB[i] = p
        rr[i] = float(t[3]) * scale
    children = [None] * n
    for ln in lines[1 + n:]:
        t = ln.split()
        if len(t) < 2:
            continue
        parent, count = int(t[0]), int(t[1])
        children[parent - 1] = [int(x) - 1 for x in t[2:2 + count]]
    return _flatten_tree(cB, rr, children, 0)


def _flatten_tree(cB, rr, children, root):
    order = [root]
    first = []
    count = []
    i = 0
    while i < len(order):
        ch = children[order[i]]
        if ch:
            first.append(len(order))
            count.append(len(ch))
            order.extend(ch)
        else:
            first.append(-1)
            count.append(0)
        i += 1
    pos = {old: new for new, old in enumerate(order)}
    rank = np.zeros(len(order), dtype=np.int32)
    k = 0
    stack = [root]
    while stack:
        nd = stack.pop()
        rank[pos[nd]] = k
        k += 1
        ch = children[nd]
        if ch:
            stack.extend(reversed(ch))
    idx = np.array(order)
    return Tree(c=cB[idx].copy(), r=rr[idx].copy(), first_child=np.array(first, dtype=np.int32),
                child_count=np.array(count, dtype=np.int32), rank=rank)


def single_sphere_tree(r):
    return Tree(c=np.zeros((1, 3)), r=np.array([float(r)]), first_child=np.array([-1], dtype=np.int32),
                child_count=np.array([0], dtype=np.int32), rank=np.array([0], dtype=np.int32))


class SceneBuilder:
    """Accumulates bodies/springs in reference parse order and emits the am3d_scene blob."""

    def __init__(self, data_root=None):
        self.bodies: list[Body] = []
        self.springs: list[SpringDef] = []
        self.trees: list[Tree] = []
        self._tree_cache = {}
        self._mesh_cache = {}
        self.data_root = data_root
        self.overrides = {}   # <collision>/<system> attributes to apply on am3d_params
        self.has_system_tag = False

    # ---- XML -------------------------------------------------------------------------------
    def parse_xml(self, path, scene=0):
        root = ET.parse(path).getroot()
        if self.data_root is None:
            # scene files reference "data/..." relative to the repository root of the reference
            self.data_root = os.path.dirname(os.path.dirname(os.path.abspath(path)))
        for el in root:
            tag = el.tag.lower()
            if tag == "system":
                self.has_system_tag = True
                for k in ("mouseSpringStiffness", "mouseSpringDamping"):
                    if k in el.attrib:
                        self.overrides[k] = float(el.attrib[k])
            elif tag == "collision":
                a = el.attrib
                if "iterations" in a:
                    self.overrides["iterations"] = int(a["iterations"])
                if "feedbackStiffness" in a:
                    self.overrides["feedback_stiffness"] = float(a["feedbackStiffness"])
                if "restitution" in a:
                    self.overrides["restitution"] = float(a["restitution"])
                    self.overrides["restitution_override"] = 1
                if "friction" in a:
                    self.overrides["friction"] = float(a["friction"])
                    self.overrides["friction_override"] = 1
                if "enablePostStabilization" in a:
                    self.overrides["enable_post_stabilization"] = int(a["enablePostStabilization"].lower() == "true")
                if "enableCompliance" in a:
                    self.overrides["enable_compliance"] = int(a["enableCompliance"].lower() == "true")
        first = len(self.bodies)
        for el in root:
            if el.tag.lower() != "body":
                continue
            t = el.attrib.get("type", "").lower()
            name = el.attrib.get("name", "")
            if t == "box":
                b = self._create_box(name, el, top=True)
            elif t == "plane":
                b = self._create_plane(name, el)
            elif t == "sphere":
                b = self._create_sphere(name, el, top=True)
            elif t == "mesh":
                b = self._create_mesh(name, el)
            elif t == "composite":
                b = self._create_composite(name, el)
            else:
                continue
            b.scene = scene
            self.bodies.append(b)
            self._resolve_springs(b, len(self.bodies) - 1, first)
        return self

    def _path(self, rel):
        return rel if os.path.isabs(rel) else os.path.join(self.data_root, rel)

    def _common(self, b: Body, el, pending_springs):
        """setCommonAttributes (XMLParser.java:546-634), tags in document order."""
        for ch in el:
            tag = ch.tag.lower()
            txt = (ch.text or "").strip()
            if tag == "x":
                b.x = np.array(_floats(txt)[:3])
            elif tag == "r":
                a = _floats(txt)
                b.R = mat_from_axis_angle(a[0], a[1], a[2], a[3])
            elif tag == "v":
                b.v = np.array(_floats(txt)[:3])
            elif tag == "omega":
                b.omega = np.array(_floats(txt)[:3])
            elif tag == "restitution":
                b.restitution = float(txt.split()[0])
            elif tag == "friction":
                b.friction = float(txt.split()[0])
            elif tag == "pinned":
                b.pinned = txt.split()[0].lower() == "true"
                if b.pinned:
                    b.minv = 0.0
                    b.jinv0 = np.zeros((3, 3))
            elif tag == "magnetic":
                b.magnetic = txt.split()[0].lower() == "true"
            elif tag == "spring":
                pending_springs.append(ch.attrib)

    def _resolve_springs(self, b, bi, first_body_of_file):
        for a in getattr(b, "_pending_springs", []):
            pB = np.array(_floats(a["pB"])[:3])
            pb1W = mat_transform(b.R, pB) + b.x
            s = None
            if "pW" in a:
                pW = np.array(_floats(a["pW"])[:3])
                d = pW - pb1W
                s = SpringDef(SPRING_WORLD, bi, -1, pB, np.zeros(3), pW, l0=math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
            elif "pB2" in a and "body2" in a:
                pB2 = np.array(_floats(a["pB2"])[:3])
                for j in range(first_body_of_file, bi):  # only bodies already in system.bodies
                    o = self.bodies[j]
                    if o.name == a["body2"]:
                        pb2W = mat_transform(o.R, pB2) + o.x
                        d = pb1W - pb2W
                        s = SpringDef(SPRING_BODYBODY, bi, j, pB, pB2, np.zeros(3),
                                      l0=math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
                        break
            else:
                s = SpringDef(SPRING_ZERO, bi, -1, pB, np.zeros(3), pb1W.copy())
            if s is None:
                continue
            if a.get("k"):
                s.k = float(a["k"])
            if a.get("d"):
                s.d = float(a["d"])
            if a.get("ls"):
                s.ls = float(a["ls"])
            self.springs.append(s)
        if hasattr(b, "_pending_springs"):
            del b._pending_springs

    def _create_box(self, name, el, top):
        s = np.array(_floats(el.attrib["dim"])[:3])
        density = float(el.attrib.get("density", 1))
        if "scale" in el.attrib:
            s = s * float(el.attrib["scale"])
        size = s.copy()
        mass = s[0] * s[1] * s[2] * density
        J = np.zeros((3, 3))
        J[0, 0] = 1.0 / 12 * mass * (s[1] * s[1] + s[2] * s[2])
        J[1, 1] = 1.0 / 12 * mass * (s[0] * s[0] + s[2] * s[2])
        J[2, 2] = 1.0 / 12 * mass * (s[0] * s[0] + s[1] * s[1])
        h = s * 0.5
        bb = [[-h[0], -h[1], -h[2]], [-h[0], -h[1], h[2]], [-h[0], h[1], -h[2]], [-h[0], h[1], h[2]],
              [h[0], -h[1], -h[2]], [h[0], -h[1], h[2]], [h[0], h[1], -h[2]], [h[0], h[1], h[2]]]
        b = _new_rigid_body(name, BODY_BOX, mass, J, False, bb)
        ps = []
        self._common(b, el, ps)
        b._pending_springs = ps
        radius = math.sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2])
        b.parts = [Part(SHAPE_BOX, size=size, radius=radius)]
        return b

    def _create_plane(self, name, el):
        p = np.array(_floats(el.attrib["p"])[:3])
        n = np.array(_floats(el.attrib["n"])[:3])
        norm = 1.0 / math.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2])
        n = n * norm
        b = _new_rigid_body(name, BODY_PLANE, 0.0, None, True, None)
        ps = []
        self._common(b, el, ps)
        b._pending_springs = ps
        b.pinned = True
        d = -(p[0] * n[0] + p[1] * n[1] + p[2] * n[2])
        b.parts = [Part(SHAPE_PLANE, size=n, radius=d, p=p)]
        return b

    def _create_sphere(self, name, el, top):
        r = float(el.attrib["r"])
        density = float(el.attrib.get("density", 1))
        mass = 4.0 / 3 * math.pi * r * r * r * density
        J = np.eye(3) * (2.0 / 5 * mass * r * r)
        bb = [[-r, -r, -r], [-r, -r, r], [-r, r, -r], [-r, r, r], [r, -r, -r], [r, -r, r], [r, r, -r], [r, r, r]]
        b = _new_rigid_body(name, BODY_SPHERE, mass, J, False, bb)
        ps = []
        self._common(b, el, ps)
        b._pending_springs = ps
        key = ("sphere", r)
        if key not in self._tree_cache:
            self.trees.append(single_sphere_tree(r))
            self._tree_cache[key] = len(self.trees) - 1
        b.parts = [Part(SHAPE_TREE, tree=self._tree_cache[key])]
        return b

    def mesh_template(self, obj, st, scale, density):
        key = (obj, st, scale, density)
        if key not in self._mesh_cache:
            mass, J, com, V = mesh_mass_properties(self._path(obj), scale, density)
            if mass < 0:
                mass = -mass
            Vc = V - com
            ll = np.minimum(Vc.min(0), _DMAX)
            ur = np.maximum(Vc.max(0), _DMIN)
            bb = [[ll[0], ll[1], ll[2]], [ll[0], ll[1], ur[2]], [ll[0], ur[1], ll[2]], [ll[0], ur[1], ur[2]],
                  [ur[0], ll[1], ll[2]], [ur[0], ll[1], ur[2]], [ur[0], ur[1], ll[2]], [ur[0], ur[1], ur[2]]]
            self.trees.append(read_sph(self._path(st), scale, com))
            self._mesh_cache[key] = (mass, J, np.array(bb), len(self.trees) - 1)
        return self._mesh_cache[key]

    def _create_mesh(self, name, el):
        scale = float(el.attrib["scale"])
        density = float(el.attrib.get("density", 1))
        mass, J, bb, tree = self.mesh_template(el.attrib["obj"], el.attrib["st"], scale, density)
        b = _new_rigid_body(name, BODY_MESH, mass, J, False, bb)
        ps = []
        self._common(b, el, ps)
        b._pending_springs = ps
        b.parts = [Part(SHAPE_TREE, tree=tree)]
        return b

    def _create_composite(self, name, el):
        subs = []
        for ch in el:
            if ch.tag.lower() != "body":
                continue
            t = ch.attrib.get("type", "").lower()
            if t == "box":
                subs.append(self._create_box(ch.attrib.get("name", ""), ch, top=False))
            elif t == "sphere":
                subs.append(self._create_sphere(ch.attrib.get("name", ""), ch, top=False))
        mass = 0.0
        com = np.zeros(3)
        for sb in subs:
            mass += sb.mass
            com = sb.mass * sb.x + com
        com = com * (1.0 / mass)
        J = np.zeros((3, 3))
        for sb in subs:
            # sub-body massAngular is R J0 R^T after its own setCommonAttributes/createBox
            J = J + (rm0rt(sb.R, sb.mass_angular0) if not sb.pinned else sb.mass_angular0 * 0)
            x, y, z = sb.x - com
            x2, y2, z2 = x * x, y * y, z * z
            op = np.array([[y2 + z2, -x * y, -x * z], [-y * x, x2 + z2, -y * z], [-z * x, -z * y, x2 + y2]])
            J = J + op * sb.mass
        ll = np.array([_DMAX] * 3)
        ur = np.array([_DMIN] * 3)
        for sb in subs:
            for p in sb.bbB:
                q = mat_transform(sb.R, p) + sb.x
                ll = np.minimum(q, ll)
                ur = np.maximum(q, ur)
        ll = ll - com
        ur = ur - com
        bb = [[ll[0], ll[1], ll[2]], [ll[0], ll[1], ur[2]], [ll[0], ur[1], ll[2]], [ll[0], ur[1], ur[2]],
              [ur[0], ur[1], ur[2]], [ur[0], ll[1], ll[2]], [ur[0], ll[1], ur[2]], [ur[0], ur[1], ll[2]]]
        b = _new_rigid_body(name, BODY_COMPOSITE, mass, J, False, bb)
        ps = []
        self._common(b, el, ps)
        b._pending_springs = ps
        b.x = b.x + com
        for sb in subs:
            part = sb.parts[0]
            part.B2C_R = sb.R.copy()
            part.B2C_t = sb.x - com
            b.parts.append(part)
        return b

    # ---- programmatic builders (synthetic configs) --------------------------------------------
    def add_plane(self, p=(0, 0, 0), n=(0, 1, 0), scene=0, name="plane"):
        el = ET.Element("body", {"p": " ".join(map(str, p)), "n": " ".join(map(str, n))})
        b = self._create_plane(name, el)
        b.scene = scene
        del b._pending_springs
        self.bodies.append(b)
        return len(self.bodies) - 1

    def add_box(self, dim, x, axis_angle=(0, 0, 1, 0), v=(0, 0, 0), omega=(0, 0, 0), pinned=False, scene=0,
                name="box", density=1.0):
        el = ET.Element("body", {"dim": " ".join(repr(float(t)) for t in dim), "density": repr(float(density))})
        ET.SubElement(el, "x").text = " ".join(repr(float(t)) for t in x)
        ET.SubElement(el, "R").text = " ".join(repr(float(t)) for t in axis_angle)
        ET.SubElement(el, "v").text = " ".join(repr(float(t)) for t in v)
        ET.SubElement(el, "omega").text = " ".join(repr(float(t)) for t in omega)
        if pinned:
            ET.SubElement(el, "pinned").text = "true"
        b = self._create_box(name, el, top=True)
        b.scene = scene
        del b._pending_springs
        self.bodies.append(b)
        return len(self.bodies) - 1

    def add_sphere(self, r, x, v=(0, 0, 0), omega=(0, 0, 0), scene=0, name="sphere"):
        el = ET.Element("body", {"r": repr(float(r))})
        ET.SubElement(el, "x").text = " ".join(repr(float(t)) for t in x)
        ET.SubElement(el, "v").text = " ".join(repr(float(t)) for t in v)
        ET.SubElement(el, "omega").text = " ".join(repr(float(t)) for t in omega)
        b = self._create_sphere(name, el, top=True)
        b.scene = scene
        del b._pending_springs
        self.bodies.append(b)
        return len(self.bodies) - 1

    def add_mesh(self, obj, st, scale, x, axis_angle=(0, 0, 1, 0), scene=0, name="mesh", density=1.0):
        mass, J, bb, tree = self.mesh_template(obj, st, scale, density)
        b = _new_rigid_body(name, BODY_MESH, mass, J, False, bb)
        b.x = np.array(x, dtype=np.float64)
        b.R = mat_from_axis_angle(*axis_angle)
        b.parts = [Part(SHAPE_TREE, tree=tree)]
        b.scene = scene
        self.bodies.append(b)
        return len(self.bodies) - 1

    # ---- emit ------------------------------------------------------------------------------
    def build(self, copies=1) -> "SceneBlob":
        """Flatten to arrays; ``copies`` > 1 replicates the whole scene as independent scene ids
        (config B: batched copies with no inter-copy interaction)."""
        nb = len(self.bodies)
        ns = sum(len(b.parts) for b in self.bodies)
        A = {}
        A["body_type"] = np.array([b.type for b in self.bodies], dtype=np.int32)
        A["body_flags"] = np.array([(F_PINNED if b.pinned else 0) | (F_MAGNETIC if b.magnetic else 0)
                                    for b in self.bodies], dtype=np.int32)
        A["body_scene"] = np.array([b.scene for b in self.bodies], dtype=np.int32)
        A["body_x"] = np.array([b.x for b in self.bodies], dtype=np.float64).reshape(nb, 3)
        A["body_R"] = np.array([b.R for b in self.bodies], dtype=np.float64).reshape(nb, 9)
        A["body_v"] = np.array([b.v for b in self.bodies], dtype=np.float64).reshape(nb, 3)
        A["body_omega"] = np.array([b.omega for b in self.bodies], dtype=np.float64).reshape(nb, 3)
        A["body_mass"] = np.array([b.mass for b in self.bodies], dtype=np.float64)
        A["body_minv"] = np.array([b.minv for b in self.bodies], dtype=np.float64)
        A["body_mass_angular0"] = np.array([b.mass_angular0 for b in self.bodies], dtype=np.float64).reshape(nb, 9)
        A["body_jinv0"] = np.array([b.jinv0 for b in self.bodies], dtype=np.float64).reshape(nb, 9)
        A["body_friction"] = np.array([b.friction for b in self.bodies], dtype=np.float64)
        A["body_restitution"] = np.array([b.restitution for b in self.bodies], dtype=np.float64)
        bb = np.zeros((nb, 8, 3))
        bbc = np.zeros(nb, dtype=np.int32)
        for i, b in enumerate(self.bodies):
            if len(b.bbB):
                bb[i] = b.bbB
                bbc[i] = 8
        A["body_bbB"] = bb.reshape(nb, 24)
        A["body_bb_count"] = bbc
        # trees → global node arrays
        tree_base = []
        nodes_c, nodes_r, nodes_fc, nodes_cc, nodes_rank = [], [], [], [], []
        base = 0
        for t in self.trees:
            tree_base.append(base)
            nodes_c.append(t.c)
            nodes_r.append(t.r)
            fc = t.first_child.copy()
            fc[fc >= 0] += base
            nodes_fc.append(fc)
            nodes_cc.append(t.child_count)
            nodes_rank.append(t.rank)
            base += len(t.r)
        A["node_c"] = np.concatenate(nodes_c).reshape(-1, 3) if nodes_c else np.zeros((0, 3))
        A["node_r"] = np.concatenate(nodes_r) if nodes_r else np.zeros(0)
        A["node_first_child"] = np.concatenate(nodes_fc).astype(np.int32) if nodes_fc else np.zeros(0, np.int32)
        A["node_child_count"] = np.concatenate(nodes_cc).astype(np.int32) if nodes_cc else np.zeros(0, np.int32)
        A["node_rank"] = np.concatenate(nodes_rank).astype(np.int32) if nodes_rank else np.zeros(0, np.int32)
        sf, sc = [], []
        st, sb, ssz, srad, sp, sR, stt, sroot = [], [], [], [], [], [], [], []
        k = 0
        for i, b in enumerate(self.bodies):
            sf.append(k)
            sc.append(len(b.parts))
            for p in b.parts:
                st.append(p.type)
                sb.append(i)
                ssz.append(p.size)
                srad.append(p.radius)
                sp.append(p.p)
                sR.append(p.B2C_R)
                stt.append(p.B2C_t)
                sroot.append(tree_base[p.tree] if p.tree >= 0 else -1)
                k += 1
        A["body_shape_first"] = np.array(sf, dtype=np.int32)
        A["body_shape_count"] = np.array(sc, dtype=np.int32)
        A["shape_type"] = np.array(st, dtype=np.int32)
        A["shape_body"] = np.array(sb, dtype=np.int32)
        A["shape_size"] = np.array(ssz, dtype=np.float64).reshape(ns, 3)
        A["shape_radius"] = np.array(srad, dtype=np.float64)
        A["shape_p"] = np.array(sp, dtype=np.float64).reshape(ns, 3)
        A["shape_B2C_R"] = np.array(sR, dtype=np.float64).reshape(ns, 9)
        A["shape_B2C_t"] = np.array(stt, dtype=np.float64).reshape(ns, 3)
        A["shape_tree_root"] = np.array(sroot, dtype=np.int32)
        nsp = len(self.springs)
        A["spring_type"] = np.array([s.type for s in self.springs], dtype=np.int32)
        A["spring_body1"] = np.array([s.body1 for s in self.springs], dtype=np.int32)
        A["spring_body2"] = np.array([s.body2 for s in self.springs], dtype=np.int32)
        A["spring_pb1"] = np.array([s.pb1 for s in self.springs], dtype=np.float64).reshape(nsp, 3)
        A["spring_pb2"] = np.array([s.pb2 for s in self.springs], dtype=np.float64).reshape(nsp, 3)
        A["spring_pw"] = np.array([s.pw for s in self.springs], dtype=np.float64).reshape(nsp, 3)
        A["spring_k"] = np.array([s.k for s in self.springs], dtype=np.float64)
        A["spring_d"] = np.array([s.d for s in self.springs], dtype=np.float64)
        A["spring_l0"] = np.array([s.l0 for s in self.springs], dtype=np.float64)
        A["spring_ls"] = np.array([s.ls for s in self.springs], dtype=np.float64)
        n_scenes = int(A["body_scene"].max()) + 1 if nb else 1
        blob = SceneBlob(A, n_scenes, names=[b.name for b in self.bodies], overrides=dict(self.overrides),
                         has_system_tag=self.has_system_tag)
        if copies > 1:
            blob = blob.replicate(copies)
        return blob


_BODY_KEYS = ["body_type", "body_flags", "body_scene", "body_shape_first", "body_shape_count", "body_x", "body_R",
              "body_v", "body_omega", "body_mass", "body_minv", "body_mass_angular0", "body_jinv0", "body_friction",
              "body_restitution", "body_bbB", "body_bb_count"]
_SHAPE_KEYS = ["shape_type", "shape_body", "shape_size", "shape_radius", "shape_p", "shape_B2C_R", "shape_B2C_t",
               "shape_tree_root"]
_NODE_KEYS = ["node_c", "node_r", "node_first_child", "node_child_count", "node_rank"]
_SPRING_KEYS = ["spring_type", "spring_body1", "spring_body2", "spring_pb1", "spring_pb2", "spring_pw", "spring_k",
                "spring_d", "spring_l0", "spring_ls"]
_I32 = {"body_type", "body_flags", "body_scene", "body_shape_first", "body_shape_count", "body_bb_count",
        "shape_type", "shape_body", "shape_tree_root", "node_first_child", "node_child_count", "node_rank",
        "spring_type", "spring_body1", "spring_body2"}


class am3d_scene(C.Structure):
    _fields_ = ([("n_bodies", C.c_int32), ("n_shapes", C.c_int32), ("n_nodes", C.c_int32), ("n_springs", C.c_int32),
                 ("n_scenes", C.c_int32), ("_pad0", C.c_int32)]
                + [(k, C.POINTER(C.c_int32) if k in _I32 else C.POINTER(C.c_double))
                   for k in ["body_type", "body_flags", "body_scene", "body_shape_first", "body_shape_count",
                             "body_x", "body_R", "body_v", "body_omega", "body_mass", "body_minv",
                             "body_mass_angular0", "body_jinv0", "body_friction", "body_restitution", "body_bbB",
                             "body_bb_count"] + _SHAPE_KEYS + _NODE_KEYS + _SPRING_KEYS])


class SceneBlob:
    """numpy arrays + a ctypes view laid out exactly as ``am3d_scene`` (include/am3d.h)."""

    def __init__(self, arrays, n_scenes=1, names=None, overrides=None, has_system_tag=False):
        self.a = {}
        for k, v in arrays.items():
            self.a[k] = np.ascontiguousarray(v, dtype=np.int32 if k in _I32 else np.float64)
        self.n_scenes = n_scenes
        self.names = names or []
        self.overrides = overrides or {}
        self.has_system_tag = has_system_tag

    @property
    def n_bodies(self):
        return len(self.a["body_type"])

    @property
    def n_shapes(self):
        return len(self.a["shape_type"])

    def body_index(self, name):
        return self.names.index(name)

    def replicate(self, copies):
        nb, ns = self.n_bodies, self.n_shapes
        nsc = self.n_scenes
        out = {}
        for k in _BODY_KEYS + _SHAPE_KEYS + _SPRING_KEYS:
            v = self.a[k]
            out[k] = np.concatenate([v] * copies) if len(v) else v
        for k in _NODE_KEYS:
            out[k] = self.a[k]
        off_b = np.repeat(np.arange(copies, dtype=np.int32) * nb, nb)
        off_s = np.repeat(np.arange(copies, dtype=np.int32) * ns, ns)
        out["body_scene"] = out["body_scene"] + np.repeat(np.arange(copies, dtype=np.int32) * nsc, nb)
        out["body_shape_first"] = out["body_shape_first"] + np.repeat(np.arange(copies, dtype=np.int32) * ns, nb)
        out["shape_body"] = out["shape_body"] + off_s // max(ns, 1) * nb
        nsp = len(self.a["spring_type"])
        if nsp:
            off = np.repeat(np.arange(copies, dtype=np.int32) * nb, nsp)
            out["spring_body1"] = out["spring_body1"] + off
            b2 = out["spring_body2"]
            out["spring_body2"] = np.where(b2 >= 0, b2 + off, b2)
        del off_b
        names = [f"{n}#{c}" for c in range(copies) for n in self.names] if copies * nb < 200000 else []
        return SceneBlob(out, nsc * copies, names=names, overrides=self.overrides, has_system_tag=self.has_system_tag)

    def as_ctypes(self):
        s = am3d_scene()
        s.n_bodies = self.n_bodies
        s.n_shapes = self.n_shapes
        s.n_nodes = len(self.a["node_r"])
        s.n_springs = len(self.a["spring_type"])
        s.n_scenes = self.n_scenes
        for k in _BODY_KEYS + _SHAPE_KEYS + _NODE_KEYS + _SPRING_KEYS:
            v = self.a[k]
            typ = C.POINTER(C.c_int32) if k in _I32 else C.POINTER(C.c_double)
            setattr(s, k, v.ctypes.data_as(typ))
        s._keepalive = self
        return s


def load_xml(path, copies=1, data_root=None) -> SceneBlob:
    return SceneBuilder(data_root=data_root).parse_xml(path).build(copies=copies)


# ---------------------------------------------------------------------------------------------
# synthetic configurations of BASELINE.json (vectorised: no per-body Python objects)
# ---------------------------------------------------------------------------------------------
class PCG32:
    """PCG32 (XSH-RR), seed 12345 wherever SURVEY.md §8d asks for jitter."""

    def __init__(self, seed=12345, seq=54):
        self.state = 0
        self.inc = ((seq << 1) | 1) & 0xFFFFFFFFFFFFFFFF
        self._next()
        self.state = (self.state + seed) & 0xFFFFFFFFFFFFFFFF
        self._next()

    def _next(self):
        old = self.state
        self.state = (old * 6364136223846793005 + self.inc) & 0xFFFFFFFFFFFFFFFF
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def uniform(self, n):
        return np.array([self._next() for _ in range(n)], dtype=np.float64) / 4294967296.0


def _boxes_blob(dims, xs, Rs, plane_y=0.0, extra_first=None):
    """plane + n boxes with per-body dims [n,3], positions [n,3], rotations [n,3,3]."""
    n = len(xs)
    nb = n + 1
    A = {}
    A["body_type"] = np.concatenate([[BODY_PLANE], np.full(n, BODY_BOX)]).astype(np.int32)
    A["body_flags"] = np.concatenate([[F_PINNED], np.zeros(n)]).astype(np.int32)
    A["body_scene"] = np.zeros(nb, np.int32)
    A["body_x"] = np.concatenate([np.zeros((1, 3)), xs])
    A["body_R"] = np.concatenate([np.eye(3).reshape(1, 9), Rs.reshape(n, 9)])
    A["body_v"] = np.zeros((nb, 3))
    A["body_omega"] = np.zeros((nb, 3))
    s = dims
    mass = s[:, 0] * s[:, 1] * s[:, 2] * 1.0
    J = np.zeros((n, 3, 3))
    J[:, 0, 0] = 1.0 / 12 * mass * (s[:, 1] * s[:, 1] + s[:, 2] * s[:, 2])
    J[:, 1, 1] = 1.0 / 12 * mass * (s[:, 0] * s[:, 0] + s[:, 2] * s[:, 2])
    J[:, 2, 2] = 1.0 / 12 * mass * (s[:, 0] * s[:, 0] + s[:, 1] * s[:, 1])
    Jinv = np.zeros((n, 3, 3))
    for d in range(3):
        Jinv[:, d, d] = 1.0 / J[:, d, d]
    A["body_mass"] = np.concatenate([[0.0], mass])
    A["body_minv"] = np.concatenate([[0.0], 1.0 / mass])
    A["body_mass_angular0"] = np.concatenate([np.zeros((1, 9)), J.reshape(n, 9)])
    A["body_jinv0"] = np.concatenate([np.zeros((1, 9)), Jinv.reshape(n, 9)])
    A["body_friction"] = np.full(nb, 0.8)
    A["body_restitution"] = np.zeros(nb)
    h = s * 0.5
    sg = np.array([[-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1], [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]],
                  dtype=np.float64)
    bb = h[:, None, :] * sg[None, :, :]
    A["body_bbB"] = np.concatenate([np.zeros((1, 24)), bb.reshape(n, 24)])
    A["body_bb_count"] = np.concatenate([[0], np.full(n, 8)]).astype(np.int32)
    A["body_shape_first"] = np.arange(nb, dtype=np.int32)
    A["body_shape_count"] = np.ones(nb, np.int32)
    A["shape_type"] = np.concatenate([[SHAPE_PLANE], np.full(n, SHAPE_BOX)]).astype(np.int32)
    A["shape_body"] = np.arange(nb, dtype=np.int32)
    A["shape_size"] = np.concatenate([[[0.0, 1.0, 0.0]], s])
    A["shape_radius"] = np.concatenate([[-(plane_y * 1.0)], np.sqrt(h[:, 0] * h[:, 0] + h[:, 1] * h[:, 1] + h[:, 2] * h[:, 2])])
    A["shape_p"] = np.concatenate([[[0.0, plane_y, 0.0]], np.zeros((n, 3))])
    A["shape_B2C_R"] = np.tile(np.eye(3).reshape(1, 9), (nb, 1))
    A["shape_B2C_t"] = np.zeros((nb, 3))
    A["shape_tree_root"] = np.full(nb, -1, np.int32)
    for k in _NODE_KEYS:
        A[k] = np.zeros((0, 3)) if k == "node_c" else np.zeros(0)
    for k in _SPRING_KEYS:
        A[k] = np.zeros((0, 3)) if k in ("spring_pb1", "spring_pb2", "spring_pw") else np.zeros(0)
    return SceneBlob(A, 1)


def box_stack(nx, ny, nz, pile=False, pitch=1.05, gap=-1e-4, seed=12345):
    """Config M: nx*nz columns × ny layers of unit boxes on the plane y=0,
    column pitch 1.05, layer gap −1e-4; ``pile`` adds U(−0.02,0.02) x/z jitter and ±0.02 rad yaw."""
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    n = len(ix)
    xs = np.empty((n, 3))
    xs[:, 0] = (ix - (nx - 1) / 2.0) * pitch
    xs[:, 2] = (iz - (nz - 1) / 2.0) * pitch
    xs[:, 1] = 0.5 + iy * (1.0 + gap) + gap
    Rs = np.tile(np.eye(3), (n, 1, 1))
    if pile:
        rng = PCG32(seed)
        u = rng.uniform(3 * n).reshape(n, 3)
        xs[:, 0] += (u[:, 0] * 2 - 1) * 0.02
        xs[:, 2] += (u[:, 1] * 2 - 1) * 0.02
        ang = (u[:, 2] * 2 - 1) * 0.02
        c, s = np.cos(ang), np.sin(ang)
        Rs[:, 0, 0] = c
        Rs[:, 0, 2] = s
        Rs[:, 2, 0] = -s
        Rs[:, 2, 2] = c
    return _boxes_blob(np.ones((n, 3)), xs, Rs)


# ---------------------------------------------------------------------------------------------
# blob persistence and instancing (fixtures for the GPU box, which has no reference checkout)
# ---------------------------------------------------------------------------------------------
def save_blob(blob: SceneBlob, path):
    np.savez_compressed(path, __n_scenes=np.array([blob.n_scenes]), __names=np.array(blob.names, dtype=object),
                        __overrides=np.array([repr(blob.overrides)], dtype=object),
                        __system=np.array([int(blob.has_system_tag)]), **blob.a)


def load_blob(path) -> SceneBlob:
    z = np.load(path, allow_pickle=True)
    arrays = {k: z[k] for k in z.files if not k.startswith("__")}
    import ast
    return SceneBlob(arrays, int(z["__n_scenes"][0]), names=list(z["__names"]),
                     overrides=ast.literal_eval(str(z["__overrides"][0])), has_system_tag=bool(z["__system"][0]))


def append_instances(blob: SceneBlob, body, xs, Rs, vs=None) -> SceneBlob:
    """Append len(xs) copies of `body` (a single-shape body, e.g. a mesh sharing its sphere tree) at the given
    poses.  Vectorised: used for the 100k sphere-tree pile of config F."""
    n = len(xs)
    a = blob.a
    nb0, ns0 = blob.n_bodies, blob.n_shapes
    assert a["body_shape_count"][body] == 1
    sh = a["body_shape_first"][body]
    out = {}
    for k in _BODY_KEYS:
        v = a[k]
        rep = np.repeat(v[body:body + 1], n, axis=0)
        out[k] = np.concatenate([v, rep])
    out["body_x"][nb0:] = xs
    out["body_R"][nb0:] = np.asarray(Rs).reshape(n, 9)
    out["body_v"][nb0:] = 0.0 if vs is None else vs
    out["body_omega"][nb0:] = 0.0
    out["body_shape_first"][nb0:] = ns0 + np.arange(n, dtype=np.int32)
    for k in _SHAPE_KEYS:
        v = a[k]
        out[k] = np.concatenate([v, np.repeat(v[sh:sh + 1], n, axis=0)])
    out["shape_body"][ns0:] = nb0 + np.arange(n, dtype=np.int32)
    for k in _NODE_KEYS + _SPRING_KEYS:
        out[k] = a[k]
    return SceneBlob(out, blob.n_scenes, names=[], overrides=blob.overrides, has_system_tag=blob.has_system_tag)


def remove_bodies(blob: SceneBlob, bodies) -> SceneBlob:
    """Drop single-shape, spring-free bodies (used to strip the template instance)."""
    a = blob.a
    keep = np.ones(blob.n_bodies, bool)
    keep[list(bodies)] = False
    keep_sh = keep[a["shape_body"]]
    new_b = np.cumsum(keep) - 1
    new_s = np.cumsum(keep_sh) - 1
    out = {}
    for k in _BODY_KEYS:
        out[k] = a[k][keep]
    for k in _SHAPE_KEYS:
        out[k] = a[k][keep_sh]
    out["body_shape_first"] = new_s[a["body_shape_first"][keep]].astype(np.int32)
    out["shape_body"] = new_b[a["shape_body"][keep_sh]].astype(np.int32)
    for k in _NODE_KEYS + _SPRING_KEYS:
        out[k] = a[k]
    if len(a["spring_type"]):
        out["spring_body1"] = new_b[a["spring_body1"]].astype(np.int32)
        b2 = a["spring_body2"]
        out["spring_body2"] = np.where(b2 >= 0, new_b[np.maximum(b2, 0)], b2).astype(np.int32)
    names = [n for n, k in zip(blob.names, keep) if k] if blob.names else []
    return SceneBlob(out, blob.n_scenes, names=names, overrides=blob.overrides, has_system_tag=blob.has_system_tag)


def random_rotations(n, seed=12345):
    """uniformly random unit axis, angle U(0, 2 pi), PCG32 stream"""
    u = PCG32(seed).uniform(4 * n).reshape(n, 4)
    z = 2 * u[:, 0] - 1
    ph = 2 * np.pi * u[:, 1]
    s = np.sqrt(np.maximum(0.0, 1 - z * z))
    ax = np.stack([s * np.cos(ph), s * np.sin(ph), z], 1)
    ang = 2 * np.pi * u[:, 2]
    c, sn = np.cos(ang), np.sin(ang)
    t = 1 - c
    x, y, zz = ax[:, 0], ax[:, 1], ax[:, 2]
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = t * x * x + c; R[:, 0, 1] = t * x * y - sn * zz; R[:, 0, 2] = t * x * zz + sn * y
    R[:, 1, 0] = t * x * y + sn * zz; R[:, 1, 1] = t * y * y + c; R[:, 1, 2] = t * y * zz - sn * x
    R[:, 2, 0] = t * x * zz - sn * y; R[:, 2, 1] = t * y * zz + sn * x; R[:, 2, 2] = t * zz * zz + c
    return R


def funnel_pile(template: SceneBlob, nx=100, ny=10, nz=100, pitch=0.95, y0=110.0, seed=12345):
    """Config F: funnel.xml without box1..3 plus nx*ny*nz instances of the torso_flux sphere-tree mesh on a lattice
    centred on x = z = 0 starting at y0.  `template` = the funnel scene with ONE mesh body as its LAST body."""
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    n = ix.size
    xs = np.empty((n, 3))
    xs[:, 0] = (ix.ravel() - (nx - 1) / 2.0) * pitch
    xs[:, 1] = y0 + iy.ravel() * pitch
    xs[:, 2] = (iz.ravel() - (nz - 1) / 2.0) * pitch
    tb = template.n_bodies - 1
    blob = append_instances(template, tb, xs, random_rotations(n, seed))
    return remove_bodies(blob, [tb])
