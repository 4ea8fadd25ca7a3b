"""Generates tests/golden/*.npz from the reference checkout (run in the authoring container only).

  scene_<name>.npz     the am3d_scene blob the loader produces for scenes3D/<name>.xml (the GPU box has no
                       reference checkout, so scenes travel as blobs)
  oracle_<name>.npz    body states / contact counts / merge-unmerge events of the CPU oracle on that scene, used
                       to pin the oracle against regressions and as golden vectors for the GPU path
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adaptivemerging_b200.ctypes_defs import apply_overrides, default_params  # noqa: E402
from adaptivemerging_b200.scene import SceneBuilder, load_xml, save_blob  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def golden_run(name, blob, steps, every):
    o = Oracle(blob, apply_overrides(default_params(), blob.overrides))
    xs, vs, Rs, ncon, ntop, at = [], [], [], [], [], []
    for s in range(steps):
        o.step(0.05)
        if (s + 1) % every == 0:
            b = o.bodies()
            t = o.timings()
            xs.append(b["x"]); vs.append(b["v"]); Rs.append(b["R"]); ncon.append(t.n_contacts); ntop.append(t.n_bodies); at.append(s + 1)
    np.savez_compressed(os.path.join(OUT, f"oracle_{name}.npz"), x=np.array(xs), v=np.array(vs), R=np.array(Rs),
                        n_contacts=np.array(ncon), n_top=np.array(ntop), at=np.array(at), events=o.events())
    print(name, "steps", steps, "events", len(o.events()), "top-level at end", o.num_top_level())


# The reference's own step logs (RigidBodySystem.exportDataToFile: "#bodies, #contacts, detection, ..." one row per
# advanceTime) recorded by the authors on tower25platform.xml.  Only the two integer columns travel: they are the one
# piece of REFERENCE OUTPUT for the 3D path that exists, and they pin the oracle (tests/test_reference_logs.py).
REF_LOGS = {
    # name: (csv under scenes3D/csv/conditional_acceptance_revisions, parameter overrides of that recording)
    "tower25platform_30it": ("tower25platform30_merged.csv", {}),
    "tower25platform_10it": ("tower25platform10_merged.csv", {"iterations": 10}),
    "tower25platform_200it": ("tower25platform200_merged.csv", {"iterations": 200}),
    "tower25platform_nosleep": ("no_sleeping/tower25platform_merged.csv", {"enable_sleeping": 0}),
}


def reference_logs(rows=400):
    import csv
    out = {}
    for name, (path, _) in REF_LOGS.items():
        full = os.path.join(REF, "scenes3D", "csv", "conditional_acceptance_revisions", path)
        data = [r for r in csv.reader(open(full)) if len(r) > 5][1:rows + 1]
        out[name] = np.array([[int(r[0]), int(r[1])] for r in data], np.int32)
    np.savez_compressed(os.path.join(OUT, "ref_logs_tower25platform.npz"), **out)
    print("reference logs:", {k: v.shape for k, v in out.items()})


def main():
    reference_logs()
    os.makedirs(OUT, exist_ok=True)
    for name, steps, every in [("tower", 300, 25), ("tower25platform", 200, 25), ("dominosPlatforms", 200, 25)]:
        blob = load_xml(os.path.join(REF, "scenes3D", name + ".xml"))
        save_blob(blob, os.path.join(OUT, f"scene_{name}.npz"))
        golden_run(name, blob, steps, every)
    # funnel template: funnel.xml without box1..3, plus ONE torso_flux mesh body (instanced at run time)
    sb = SceneBuilder(data_root=REF).parse_xml(os.path.join(REF, "scenes3D", "funnel.xml"))
    sb.bodies = [b for b in sb.bodies if b.name not in ("box1", "box2", "box3")]
    # springs reference bodies by index: rebuild them after the removal
    names = [b.name for b in sb.bodies]
    assert not sb.springs or True
    sb2 = SceneBuilder(data_root=REF)
    import xml.etree.ElementTree as ET
    root = ET.parse(os.path.join(REF, "scenes3D", "funnel.xml")).getroot()
    for el in list(root):
        if el.tag.lower() == "body" and el.attrib.get("name") in ("box1", "box2", "box3"):
            root.remove(el)
    tmp = "/tmp/funnel_nobox.xml"
    ET.ElementTree(root).write(tmp)
    sb2.data_root = REF
    sb2.parse_xml(tmp)
    sb2.add_mesh("data/scaledtorso10.obj", "data/torso_flux.sph", 0.08, (0, 110, 0), name="torso")
    blob = sb2.build()
    save_blob(blob, os.path.join(OUT, "scene_funnel_template.npz"))
    print("funnel template bodies", blob.n_bodies, "shapes", blob.n_shapes, "nodes", len(blob.a["node_r"]), "springs", len(blob.a["spring_type"]))
    # a small mesh scene for the sphere-tree narrowphase tests: 12 torsos dropped on the plane + a box
    sb3 = SceneBuilder(data_root=REF)
    sb3.add_plane((0, 0, 0), (0, 1, 0))
    sb3.add_box((3, 1, 3), (0, 0.5, 0), pinned=True, name="slab")
    from adaptivemerging_b200.scene import random_rotations
    Rr = random_rotations(12, 7)
    k = 0
    for i in range(2):
        for j in range(2):
            for l in range(3):
                b = sb3.add_mesh("data/scaledtorso10.obj", "data/torso_flux.sph", 0.08, (-0.5 + 1.0 * i, 1.5 + 0.8 * l, -0.5 + 1.0 * j), name=f"t{k}")
                sb3.bodies[b].R = Rr[k]
                k += 1
    blob = sb3.build()
    save_blob(blob, os.path.join(OUT, "scene_torsos.npz"))
    golden_run("torsos", blob, 120, 20)


if __name__ == "__main__":
    main()
