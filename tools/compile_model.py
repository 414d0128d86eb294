#!/usr/bin/env python
"""Model compiler: reference MuJoCo XML + STL hulls -> flat SoA model tables (npz).

Reads (in THIS container only) the reference's frozen neutral-shape humanoid
    /root/reference/assets/mujoco_models/humanoid_smpl_neutral_mesh.xml   (tree, joints, actuators)
    /root/reference/assets/mujoco_models/geom/<Body>.stl                  (24 convex body hulls)
and writes   uhc_b200/assets/smpl_neutral_model.npz   -- numeric tables only, no reference source.

What is derived and how (Appendix A/B):
  * coordinate="global" (xml:2): body pos are world coordinates in the rest pose, every quat is identity,
    so the parent-relative offset is a plain difference; joint anchors coincide with the body origin
    (xml:52-54), three hinges per non-root body with axes z,y,x; one free joint on the Pelvis.
  * inertiafromgeom="true" (xml:2) + mesh geoms at density 1000: mass, centre of mass and inertia tensor are
    exact polyhedral integrals over the STL triangles (signed tetrahedra against the origin).
  * joint armature 0.01 on hinges / 0 on the root (xml:9,51), damping = stiffness = 0.
  * floor contact only: body geoms contype 0 / conaffinity 1, floor contype 7 (default) / conaffinity 1
    (xml:10,50,53) -> body-body pairs never collide, floor-body always does. condim = max(3,1) = 3,
    margin = max(0.001, 0.001), friction = (1, .005, .0001) -> mu = 1; solref/solimp defaults.
  * hull graph: scipy ConvexHull over the unique STL vertices -> hull vertices + vertex adjacency
    (ascending neighbour order), used by the plane/convex manifold rule (oracle/uhc_oracle.c).
  * body_invweight0: (trace of J M^-1 J^T at qpos0)/3 for translation at the body's centre of mass.
  * PD gains / torque limits / diff weights: the SMPLConverter tables, uhc/smpllib/smpl_mujoco.py:40-91,
    replicated x3 per body as in :271-281.
"""
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = os.environ.get("UHC_REFERENCE", "/root/reference")
XML = os.path.join(REF, "assets/mujoco_models/humanoid_smpl_neutral_mesh.xml")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "uhc_b200", "assets", "smpl_neutral_model.npz")

# uhc/smpllib/smpl_mujoco.py:40-65 (diff weights) and :67-91 (kp, kd, gear, torque limit)
BODY_WS = {"Pelvis": 1, "L_Hip": 1, "L_Knee": 1, "L_Ankle": 1, "L_Toe": 0, "R_Hip": 1, "R_Knee": 1, "R_Ankle": 1,
           "R_Toe": 0, "Torso": 1, "Spine": 1, "Chest": 1, "Neck": 1, "Head": 1, "L_Thorax": 1, "L_Shoulder": 1,
           "L_Elbow": 1, "L_Wrist": 1, "L_Hand": 0, "R_Thorax": 1, "R_Shoulder": 1, "R_Elbow": 1, "R_Wrist": 1,
           "R_Hand": 0}
BODY_PARAMS = {"L_Hip": (500, 50, 500), "L_Knee": (500, 50, 500), "L_Ankle": (400, 40, 500), "L_Toe": (200, 20, 500),
               "R_Hip": (500, 50, 500), "R_Knee": (500, 50, 500), "R_Ankle": (400, 40, 500), "R_Toe": (200, 20, 500),
               "Torso": (1000, 100, 500), "Spine": (1000, 100, 500), "Chest": (1000, 100, 500),
               "Neck": (100, 10, 250), "Head": (100, 10, 250), "L_Thorax": (400, 40, 500),
               "L_Shoulder": (400, 40, 500), "L_Elbow": (300, 30, 150), "L_Wrist": (100, 10, 150),
               "L_Hand": (100, 10, 150), "R_Thorax": (400, 40, 150), "R_Shoulder": (400, 40, 250),
               "R_Elbow": (300, 30, 150), "R_Wrist": (100, 10, 150), "R_Hand": (100, 10, 150)}
EE_NAMES = ["L_Ankle", "R_Ankle", "L_Wrist", "R_Wrist", "Head"]  # uhc/smpllib/smpl_parser.py:228


def read_stl(path):
    raw = open(path, "rb").read()
    n = struct.unpack("<I", raw[80:84])[0]
    assert len(raw) == 84 + 50 * n, "binary STL expected"
    rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
    return rec["v"].astype(np.float64)  # (n,3,3)


def poly_mass_props(tri):
    """Exact volume, COM, inertia-about-COM (density 1) of a closed triangle mesh."""
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    det = np.einsum("ij,ij->i", a, np.cross(b, c))  # 6 * signed tet volume
    vol = det.sum() / 6.0
    sgn = 1.0 if vol > 0 else -1.0
    det = det * sgn
    vol = abs(vol)
    com = (det[:, None] * (a + b + c)).sum(0) / (24.0 * vol)
    # second moments  int x_i x_j dV  over tets (origin,a,b,c)
    S = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            t = (2 * a[:, i] * a[:, j] + 2 * b[:, i] * b[:, j] + 2 * c[:, i] * c[:, j]
                 + a[:, i] * b[:, j] + a[:, j] * b[:, i] + a[:, i] * c[:, j] + a[:, j] * c[:, i]
                 + b[:, i] * c[:, j] + b[:, j] * c[:, i])
            S[i, j] = (det * t).sum() / 120.0
    S -= vol * np.outer(com, com)  # central second moments
    I = np.trace(S) * np.eye(3) - S
    return vol, com, I


def main():
    from scipy.spatial import ConvexHull

    root = ET.parse(XML).getroot()
    names, parent, gpos = [], [], []
    jrange = []          # hinge ranges in radians, dof order (xml default: limited="true", angle unit degree)

    def walk(el, par):
        for b in el.findall("body"):
            idx = len(names)
            names.append(b.get("name"))
            parent.append(par)
            gpos.append([float(x) for x in b.get("pos").split()])
            js = b.findall("joint")
            if par < 0:
                assert len(js) == 1 and js[0].get("type") == "free"
            else:
                assert [j.get("axis").split()[k] for k, j in zip((2, 1, 0), js)] == ["1.0000"] * 3, "hinge order z,y,x"
                for j in js:
                    assert j.get("limited", "true") == "true"
                    jrange.append([np.deg2rad(float(x)) for x in j.get("range").split()])
            walk(b, idx)

    walk(root.find("worldbody"), -1)
    nb = len(names)
    assert nb == 24
    gpos = np.array(gpos)
    parent = np.array(parent, dtype=np.int32)
    offset = gpos.copy()
    offset[1:] = gpos[1:] - gpos[parent[1:]]
    dt = float(root.find("option").get("timestep"))
    acts = [m.get("joint") for m in root.find("actuator").findall("motor")]
    assert len(acts) == 69 and acts[0] == names[1] + "_z"

    mass = np.zeros(nb)
    ipos = np.zeros((nb, 3))
    inertia = np.zeros((nb, 3, 3))
    verts, vadr, vnum, nbr, nbradr = [], [], [], [], [0]
    for b, nm in enumerate(names):
        tri = read_stl(os.path.join(REF, "assets/mujoco_models/geom", nm + ".stl"))
        vol, com, I = poly_mass_props(tri)
        mass[b] = 1000.0 * vol
        ipos[b] = com - gpos[b]
        inertia[b] = 1000.0 * I
        uv = np.unique(tri.reshape(-1, 3), axis=0)
        hull = ConvexHull(uv)
        hv = np.sort(hull.vertices)
        remap = -np.ones(len(uv), dtype=np.int64)
        remap[hv] = np.arange(len(hv))
        adj = [set() for _ in hv]
        for s in hull.simplices:
            r = remap[s]
            for i in range(3):
                for j in range(3):
                    if i != j:
                        adj[r[i]].add(int(r[j]))
        vadr.append(sum(vnum))
        vnum.append(len(hv))
        verts.append(uv[hv] - gpos[b])  # body-local frame (rest rotation = identity)
        for a in adj:
            nbr.extend(sorted(a))
            nbradr.append(len(nbr))
    verts = np.concatenate(verts)

    # dof tables
    nv, nq, nu = 6 + 3 * (nb - 1), 7 + 3 * (nb - 1), 3 * (nb - 1)
    dof_body = np.array([0] * 6 + [1 + (d // 3) for d in range(nu)], dtype=np.int32)
    armature = np.array([0.0] * 6 + [0.01] * nu)
    jkp = np.concatenate([[BODY_PARAMS[n][0]] * 3 for n in names[1:]]).astype(np.float64)
    jkd = np.concatenate([[BODY_PARAMS[n][1]] * 3 for n in names[1:]]).astype(np.float64)
    tlim = np.concatenate([[BODY_PARAMS[n][2]] * 3 for n in names[1:]]).astype(np.float64)
    diffw = np.array([BODY_WS[n] for n in names], dtype=np.float64)

    # M at qpos0 (all rotations identity, axes world aligned) by the Jacobian method -> body_invweight0
    axes = np.eye(3)[[2, 1, 0]]  # z, y, x
    anc = []
    for b in range(nb):
        ch, p = [], b
        while p >= 0:
            ch.append(p)
            p = parent[p]
        anc.append(ch)
    xipos = gpos + ipos
    Jv = np.zeros((nb, 3, nv))
    Jw = np.zeros((nb, 3, nv))
    for b in range(nb):
        Jv[b, :, 0:3] = np.eye(3)
        for k in range(3):
            Jw[b, :, 3 + k] = np.eye(3)[k]
            Jv[b, :, 3 + k] = np.cross(np.eye(3)[k], xipos[b] - gpos[0])
        for a in anc[b]:
            if a == 0:
                continue
            for k in range(3):
                d = 6 + 3 * (a - 1) + k
                Jw[b, :, d] = axes[k]
                Jv[b, :, d] = np.cross(axes[k], xipos[b] - gpos[a])
    M = np.diag(armature)
    for b in range(nb):
        M += mass[b] * Jv[b].T @ Jv[b] + Jw[b].T @ inertia[b] @ Jw[b]
    Minv = np.linalg.inv(M)
    invw = np.zeros((nb, 2))
    for b in range(nb):
        invw[b, 0] = np.trace(Jv[b] @ Minv @ Jv[b].T) / 3
        invw[b, 1] = np.trace(Jw[b] @ Minv @ Jw[b].T) / 3
    # dof_invweight0 (diagApprox of the joint-limit rows): diagonal of M^-1 at qpos0; the free joint's translational / rotational dofs share their means
    dof_invw = np.diag(Minv).copy()
    dof_invw[0:3], dof_invw[3:6] = dof_invw[0:3].mean(), dof_invw[3:6].mean()

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(
        OUT, body_names=np.array(names), parent=parent, body_gpos=gpos, body_offset=offset, body_mass=mass,
        body_ipos=ipos, body_inertia=inertia, body_invweight0=invw, dof_body=dof_body, armature=armature,
        jkp=jkp, jkd=jkd, torque_lim=tlim, diffw=diffw, jnt_range=np.array(jrange), dof_invweight0=dof_invw,
        ee_body=np.array([names.index(n) for n in EE_NAMES], dtype=np.int32),
        hull_vert=verts, hull_vadr=np.array(vadr, dtype=np.int32), hull_vnum=np.array(vnum, dtype=np.int32),
        hull_nbr=np.array(nbr, dtype=np.int32), hull_nbradr=np.array(nbradr, dtype=np.int32),
        timestep=np.float64(dt), frame_skip=np.int32(15), gravity=np.array([0, 0, -9.81]),
        margin=np.float64(0.001), friction=np.float64(1.0), solref=np.array([0.02, 1.0]),
        solimp=np.array([0.9, 0.95, 0.001, 0.5, 2.0]), nq=np.int32(nq), nv=np.int32(nv), nu=np.int32(nu))
    print("bodies", nb, "nq/nv/nu", nq, nv, nu, "mass %.3f kg" % mass.sum(), "hull verts", len(verts),
          "max verts/body", max(vnum), "max nbrs", max(np.diff(nbradr)), "dt", dt)
    print("invweight0 tran: root %.5f foot %.5f hand %.5f" % (invw[0, 0], invw[3, 0], invw[18, 0]))


if __name__ == "__main__":
    sys.exit(main())
