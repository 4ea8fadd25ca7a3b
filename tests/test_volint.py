"""CPU: the mesh mass properties of the loader mirror (adaptivemerging_b200/volint.py, restating tools/moments/VolInt.java,
Polyhedron.java, PolygonSoup.java) against closed forms and against an independent formulation (signed tetrahedra)."""
import os

import numpy as np
import pytest

from adaptivemerging_b200 import volint
from tests.conftest import REF

CUBE_V = np.array([[x, y, z] for x in (0., 1.) for y in (0., 1.) for z in (0., 1.)])
# outward-facing triangles of the unit cube (vertex index = 4x + 2y + z)
CUBE_F = [[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]]


def tetra_sums(V, faces, density):
    """Independent formulation: signed tetrahedra (origin, a, b, c)."""
    F = np.array([f[:3] for f in faces])
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    det = np.einsum("ij,ij->i", a, np.cross(b, c))
    vol = det.sum() / 6.0
    com = ((a + b + c) * det[:, None]).sum(0) / (24.0 * vol)
    S = a + b + c
    cov = sum(np.einsum("n,ni,nj->ij", det, p, p) for p in (a, b, c, S)) / 120.0
    J = density * (np.trace(cov) * np.eye(3) - cov)
    mass = density * vol
    J = J - mass * ((com @ com) * np.eye(3) - np.outer(com, com))
    return mass, J, com


def test_unit_cube_closed_form():
    mass, J, com, ndeg = volint.compute_mass_properties(CUBE_V, CUBE_F, 3.0)
    assert ndeg == 0
    assert abs(mass - 3.0) < 1e-15
    assert np.abs(com - 0.5).max() < 1e-15
    assert np.abs(J - np.eye(3) * 3.0 / 6.0).max() < 1e-15


def test_box_moved_scaled_rotated():
    rng = np.random.default_rng(5)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    dim = np.array([2.0, 0.5, 1.25])
    t = np.array([3.0, -7.0, 0.4])
    V = ((CUBE_V - 0.5) * dim) @ R.T + t
    mass, J, com, _ = volint.compute_mass_properties(V, CUBE_F, 2.0)
    m = 2.0 * dim.prod()
    Jb = np.diag([m / 12 * (dim[1] ** 2 + dim[2] ** 2), m / 12 * (dim[0] ** 2 + dim[2] ** 2), m / 12 * (dim[0] ** 2 + dim[1] ** 2)])
    assert abs(mass - m) < 1e-12
    assert np.abs(com - t).max() < 1e-12
    assert np.abs(J - R @ Jb @ R.T).max() < 1e-11
    assert np.array_equal(J, J.T)  # the reference stores one value in both triangles


def test_degenerate_faces_are_dropped_and_polygons_use_three_vertices(tmp_path):
    faces = [f + [f[0]] for f in CUBE_F] + [[0, 0, 5], [2, 2, 2]]  # a trailing 4th index is ignored (PolygonSoup.getPolyhedron)
    m0 = volint.compute_mass_properties(CUBE_V, CUBE_F, 1.0)
    m1 = volint.compute_mass_properties(CUBE_V, faces, 1.0)
    assert m1[3] == 2
    assert m0[0] == m1[0] and np.array_equal(m0[1], m1[1]) and np.array_equal(m0[2], m1[2])
    # OBJ reader: "v " and "f " lines only, v/t/n tuples keep the vertex index, 1-based
    p = tmp_path / "cube.obj"
    with open(p, "w") as f:
        f.write("# cube\nvn 0 0 1\n")
        for v in CUBE_V:
            f.write(f"v {v[0]} {v[1]} {v[2]}\n")
        for fc in CUBE_F:
            f.write("f " + " ".join(f"{i + 1}/1/1" for i in fc) + "\n")
    mass, J, com, V = volint.mesh_mass_properties(str(p), 2.0, 1.0)
    assert abs(mass - 8.0) < 1e-14 and np.abs(com - 1.0).max() < 1e-15
    assert np.array_equal(V, CUBE_V * 2.0)


def _blob_mesh(n=24):
    """A closed, bumpy, star-shaped surface (lat/long grid), off-centre."""
    th = np.linspace(0, np.pi, n + 1)[1:-1]
    ph = np.linspace(0, 2 * np.pi, 2 * n, endpoint=False)
    r = lambda t, p: 1.0 + 0.3 * np.sin(3 * t) * np.cos(2 * p) + 0.1 * np.cos(5 * p)  # noqa: E731
    V = [[0, 0, r(0, 0)]]
    for t in th:
        for p in ph:
            V.append([r(t, p) * np.sin(t) * np.cos(p), r(t, p) * np.sin(t) * np.sin(p), r(t, p) * np.cos(t)])
    V.append([0, 0, -r(np.pi, 0)])
    V = np.array(V) + np.array([0.3, -0.2, 0.7])
    m = 2 * n
    idx = lambda i, j: 1 + i * m + (j % m)  # noqa: E731
    F = []
    for j in range(m):
        F.append([0, idx(0, j), idx(0, j + 1)])
        F.append([len(V) - 1, idx(n - 2, j + 1), idx(n - 2, j)])
    for i in range(n - 2):
        for j in range(m):
            F.append([idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)])
            F.append([idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)])
    return V, F


def test_against_signed_tetrahedra_on_a_curved_mesh():
    V, F = _blob_mesh()
    mass, J, com, ndeg = volint.compute_mass_properties(V, F, 1.7)
    m2, J2, c2 = tetra_sums(V, F, 1.7)
    assert ndeg == 0 and mass > 0
    assert abs(mass - m2) < 1e-12 * mass
    assert np.abs(com - c2).max() < 1e-12
    assert np.abs(J - J2).max() < 1e-11 * np.abs(J).max()
    assert np.all(np.linalg.eigvalsh(J) > 0)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present")
def test_torso_mesh_of_the_reference_and_the_committed_blobs():
    """scaledtorso10.obj (5 672 faces) at the scale the funnel scenes use; the committed scene blobs carry these doubles."""
    from tests.util import golden_scene
    mass, J, com, V = volint.mesh_mass_properties(os.path.join(REF, "data", "scaledtorso10.obj"), 0.08, 1.0)
    faces = volint.read_obj(os.path.join(REF, "data", "scaledtorso10.obj"))[1]
    m2, J2, c2 = tetra_sums(V, faces, 1.0)
    assert abs(mass - m2) < 1e-13 * mass and np.abs(com - c2).max() < 1e-13 and np.abs(J - J2).max() < 1e-12 * np.abs(J).max()
    assert mass in golden_scene("torsos").a["body_mass"].tolist()
