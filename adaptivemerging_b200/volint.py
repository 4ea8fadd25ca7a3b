"""Mass properties of a closed triangle mesh with the arithmetic of the reference's loader.

Host-side mirror (load time, caller side of the boundary) of
  tools/moments/PolygonSoup.java:38-47, 80-115, 157-166   OBJ reader: "v " / "f " lines, first index of every v/t/n tuple,
                                                           the first THREE vertices of a face
  tools/moments/Polyhedron.java:216-238                   face normal, offset w, degenerate faces dropped
  tools/moments/VolInt.java:105-181                       projection integrals of a face (Mirtich 1996)
  tools/moments/VolInt.java:183-227                       face integrals
  tools/moments/VolInt.java:229-275                       volume integrals, accumulated face after face
  tools/moments/VolInt.java:297-359                       mass, centre of mass, inertia about the centre of mass

Every expression keeps the reference's operand order and association, so that a scene loaded here carries the same
doubles the Java loader produces (IEEE arithmetic is deterministic; numpy evaluates the element-wise operations below
exactly as written, one rounding per operation).  Work that is independent per face is vectorised over the faces;
the sums over the faces are taken one face after the other, as the reference's loop does.
"""
import numpy as np


def read_obj(path):
    """PolygonSoup(String) :38-47, parseVertex :80-89, parseFace :97-115."""
    verts, faces = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                t = line[2:].split()
                verts.append((float(t[0]), float(t[1]), float(t[2])))
            elif line.startswith("f "):
                faces.append([int(tok.split("/")[0]) - 1 for tok in line[2:].split()])
    return np.array(verts, np.float64).reshape(-1, 3), faces


def _faces(V, faces):
    """Polyhedron.addFace(Point3d x3) + addFaceHelper :216-238: vertices, unit normal, w of the non-degenerate faces."""
    F = np.array([f[:3] for f in faces], np.int64).reshape(-1, 3)
    P = V[F]  # [nf, 3 vertices, 3 coordinates]
    d1 = P[:, 1] - P[:, 0]
    d2 = P[:, 2] - P[:, 1]
    nx = d1[:, 1] * d2[:, 2] - d2[:, 1] * d1[:, 2]
    ny = d1[:, 2] * d2[:, 0] - d2[:, 2] * d1[:, 0]
    nz = d1[:, 0] * d2[:, 1] - d2[:, 0] * d1[:, 1]
    ln = np.sqrt(nx * nx + ny * ny + nz * nz)
    keep = ln != 0
    P, nx, ny, nz, ln = P[keep], nx[keep], ny[keep], nz[keep], ln[keep]
    n = np.stack([nx / ln, ny / ln, nz / ln], 1)
    w = -n[:, 0] * P[:, 0, 0] - n[:, 1] * P[:, 0, 1] - n[:, 2] * P[:, 0, 2]
    return P, n, w, int((~keep).sum())


def _pick(a, idx):
    return np.take_along_axis(a, idx[:, None], 1)[:, 0]


def volume_integrals(P, n, w):
    """compVolumeIntegrals :229-275 (with compFaceIntegrals :183-227 and compProjectionIntegrals :105-181 inlined per
    face).  Returns T0, T1[3], T2[3], TP[3]."""
    nf = len(P)
    ax, ay, az = np.abs(n[:, 0]), np.abs(n[:, 1]), np.abs(n[:, 2])
    C = np.where((ax > ay) & (ax > az), 0, np.where(ay > az, 1, 2))
    A = (C + 1) % 3
    B = (A + 1) % 3
    # --- projection integrals: the three edges in turn, every accumulator starts at 0.0 ---
    z = np.zeros(nf)
    P1, Pa, Pb, Paa, Pab, Pbb, Paaa, Paab, Pabb, Pbbb = (z.copy() for _ in range(10))
    for i in range(3):
        j = (i + 1) % 3
        a0, b0 = _pick(P[:, i], A), _pick(P[:, i], B)
        a1, b1 = _pick(P[:, j], A), _pick(P[:, j], B)
        da = a1 - a0
        db = b1 - b0
        a0_2 = a0 * a0; a0_3 = a0_2 * a0; a0_4 = a0_3 * a0
        b0_2 = b0 * b0; b0_3 = b0_2 * b0; b0_4 = b0_3 * b0
        a1_2 = a1 * a1; a1_3 = a1_2 * a1
        b1_2 = b1 * b1; b1_3 = b1_2 * b1
        C1 = a1 + a0
        Ca = a1 * C1 + a0_2
        Caa = a1 * Ca + a0_3
        Caaa = a1 * Caa + a0_4
        Cb = b1 * (b1 + b0) + b0_2
        Cbb = b1 * Cb + b0_3
        Cbbb = b1 * Cbb + b0_4
        Cab = 3 * a1_2 + 2 * a1 * a0 + a0_2
        Kab = a1_2 + 2 * a1 * a0 + 3 * a0_2
        Caab = a0 * Cab + 4 * a1_3
        Kaab = a1 * Kab + 4 * a0_3
        Cabb = 4 * b1_3 + 3 * b1_2 * b0 + 2 * b1 * b0_2 + b0_3
        Kabb = b1_3 + 2 * b1_2 * b0 + 3 * b1 * b0_2 + 4 * b0_3
        P1 = P1 + db * C1
        Pa = Pa + db * Ca
        Paa = Paa + db * Caa
        Paaa = Paaa + db * Caaa
        Pb = Pb + da * Cb
        Pbb = Pbb + da * Cbb
        Pbbb = Pbbb + da * Cbbb
        Pab = Pab + db * (b1 * Cab + b0 * Kab)
        Paab = Paab + db * (b1 * Caab + b0 * Kaab)
        Pabb = Pabb + da * (a1 * Cabb + a0 * Kabb)
    P1 = P1 / 2.0; Pa = Pa / 6.0; Paa = Paa / 12.0; Paaa = Paaa / 20.0
    Pb = Pb / -6.0; Pbb = Pbb / -12.0; Pbbb = Pbbb / -20.0
    Pab = Pab / 24.0; Paab = Paab / 60.0; Pabb = Pabb / -60.0
    # --- face integrals ---
    nA, nB, nC = _pick(n, A), _pick(n, B), _pick(n, C)
    k1 = 1 / nC
    k2 = k1 * k1
    k3 = k2 * k1
    k4 = k3 * k1
    sq = lambda x: x * x            # noqa: E731  VolInt.SQR
    cube = lambda x: x * x * x      # noqa: E731  VolInt.CUBE
    Fa = k1 * Pa
    Fb = k1 * Pb
    Fc = -k2 * (nA * Pa + nB * Pb + w * P1)
    Faa = k1 * Paa
    Fbb = k1 * Pbb
    Fcc = k3 * (sq(nA) * Paa + 2 * nA * nB * Pab + sq(nB) * Pbb + w * (2 * (nA * Pa + nB * Pb) + w * P1))
    Faaa = k1 * Paaa
    Fbbb = k1 * Pbbb
    Fccc = -k4 * (cube(nA) * Paaa + 3 * sq(nA) * nB * Paab + 3 * nA * sq(nB) * Pabb + cube(nB) * Pbbb
                  + 3 * w * (sq(nA) * Paa + 2 * nA * nB * Pab + sq(nB) * Pbb) + w * w * (3 * (nA * Pa + nB * Pb) + w * P1))
    Faab = k1 * Paab
    Fbbc = -k2 * (nA * Pabb + nB * Pbbb + w * Pbb)
    Fcca = k3 * (sq(nA) * Paaa + 2 * nA * nB * Paab + sq(nB) * Pabb + w * (2 * (nA * Paa + nB * Pab) + w * Pa))
    # --- per-face terms of the volume integrals, then the sums face after face ---
    t0 = n[:, 0] * np.where(A == 0, Fa, np.where(B == 0, Fb, Fc))
    terms = {"T1": (nA * Faa, nB * Fbb, nC * Fcc), "T2": (nA * Faaa, nB * Fbbb, nC * Fccc), "TP": (nA * Faab, nB * Fbbc, nC * Fcca)}
    T0 = 0.0
    T = {k: [0.0, 0.0, 0.0] for k in terms}
    Al, Bl, Cl = A.tolist(), B.tolist(), C.tolist()
    t0l = t0.tolist()
    tl = {k: tuple(x.tolist() for x in v) for k, v in terms.items()}
    for f in range(nf):
        T0 += t0l[f]
        a, b, c = Al[f], Bl[f], Cl[f]
        for k in ("T1", "T2", "TP"):
            ta, tb, tc = tl[k]
            T[k][a] += ta[f]
            T[k][b] += tb[f]
            T[k][c] += tc[f]
    T1 = [x / 2 for x in T["T1"]]
    T2 = [x / 3 for x in T["T2"]]
    TP = [x / 2 for x in T["TP"]]
    return T0, T1, T2, TP


def compute_mass_properties(V, faces, density):
    """VolInt.computeMassProperties :297-359 -> (mass, inertia about the centre of mass [3,3], centre of mass [3],
    number of degenerate faces dropped)."""
    P, n, w, ndeg = _faces(V, faces)
    T0, T1, T2, TP = volume_integrals(P, n, w)
    X, Y, Z = 0, 1, 2
    mass = density * T0
    r = [T1[X] / T0, T1[Y] / T0, T1[Z] / T0]
    J = [[0.0] * 3 for _ in range(3)]
    J[X][X] = density * (T2[Y] + T2[Z])
    J[Y][Y] = density * (T2[Z] + T2[X])
    J[Z][Z] = density * (T2[X] + T2[Y])
    J[X][Y] = J[Y][X] = -density * TP[X]
    J[Y][Z] = J[Z][Y] = -density * TP[Y]
    J[Z][X] = J[X][Z] = -density * TP[Z]
    J[X][X] -= mass * (r[Y] * r[Y] + r[Z] * r[Z])
    J[Y][Y] -= mass * (r[Z] * r[Z] + r[X] * r[X])
    J[Z][Z] -= mass * (r[X] * r[X] + r[Y] * r[Y])
    # Java: J[X][Y] = J[Y][X] += e  evaluates J[Y][X] + e once and stores it in both
    J[X][Y] = J[Y][X] = J[Y][X] + mass * r[X] * r[Y]
    J[Y][Z] = J[Z][Y] = J[Z][Y] + mass * r[Y] * r[Z]
    J[Z][X] = J[X][Z] = J[X][Z] + mass * r[Z] * r[X]
    return mass, np.array(J), np.array(r), ndeg


def mesh_mass_properties(obj_path, scale, density):
    """XMLParser.createMesh :457-464: vertices scaled (Tuple3d.scale: one multiplication per coordinate), then VolInt.
    Returns (mass, inertia, centre of mass, scaled vertices); the caller flips a negative mass and recentres."""
    V, faces = read_obj(obj_path)
    V = V * scale
    mass, J, com, _ = compute_mass_properties(V, faces, density)
    return mass, J, com, V
