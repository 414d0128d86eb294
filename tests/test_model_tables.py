"""CPU: invariants of the model tables the kernels walk (uhc_b200/model.py) -- in particular the centre-rooted elimination tree of
the articulated-body solve (Model::lvl_pack): every body once, consistent parent / child groups, every joint block used exactly once,
reversed joints exactly on the path centre -> Pelvis."""
import numpy as np

from uhc_b200.model import HumanoidModel, NB, NV


def _decode(e):
    return dict(body=(e & 63) - 1, pgrp=(e >> 6) & 7, ch=[((e >> (9 + 3 * k)) & 7) - 1 for k in range(3)], nslot=(e >> 18) & 3,
                jblk=(e >> 20) & 31, rev=(e >> 25) & 1, nlvl=(e >> 26) & 15)


def test_solve_tree_table():
    m = HumanoidModel()
    t = m.lvl_pack.reshape(9, 5)
    nlvl = m.solve_levels
    assert 1 <= nlvl <= 9 and all(_decode(int(e))["nlvl"] == nlvl for e in t.reshape(-1))
    ent = [[_decode(int(e)) for e in row] for row in t]
    bodies = [d["body"] for row in ent for d in row if d["body"] >= 0]
    assert sorted(bodies) == list(range(NB))                                  # every body exactly once
    assert all(d["body"] < 0 for row in ent[nlvl:] for d in row)               # nothing beyond the last level
    assert [d["body"] for d in ent[0] if d["body"] >= 0] == [m.solve_root]     # one centre body
    # the centre minimises the tree height
    adj = [[c for c in range(NB) if m.parent[c] == b] + ([int(m.parent[b])] if b > 0 else []) for b in range(NB)]

    def height(r):
        d, order = {r: 0}, [r]
        for b in order:
            for c in adj[b]:
                if c not in d:
                    d[c] = d[b] + 1
                    order.append(c)
        return max(d.values())
    assert height(m.solve_root) + 1 == nlvl == min(height(r) for r in range(NB)) + 1
    blocks, rev = [], []
    for L in range(nlvl):
        for g, d in enumerate(ent[L]):
            if d["body"] < 0:
                continue
            assert d["nslot"] == max(sum(c >= 0 for c in x["ch"]) for x in ent[L] if x["body"] >= 0)
            for c in d["ch"]:
                if c >= 0:                                                   # child group sits one level deeper and points back here
                    child = ent[L + 1][c]
                    assert child["body"] >= 0 and child["pgrp"] == g
                    assert m.parent[child["body"]] == d["body"] or m.parent[d["body"]] == child["body"]
            if L == 0:
                assert d["jblk"] == NV // 3 and d["rev"] == 0                  # virtual dofs NV .. NV+5
            else:
                blocks.append(d["jblk"])
                owner = d["jblk"] - 1                                         # body whose kinematic joint this block is
                assert owner == (m.parent[ent[L - 1][d["pgrp"]]["body"]] == d["body"] and ent[L - 1][d["pgrp"]]["body"] or d["body"])
                if d["rev"]:
                    rev.append(d["body"])
    assert sorted(blocks) == list(range(2, NB + 1))                           # joint blocks of bodies 1..23, each once
    path, b = [], m.solve_root                                                # reversed joints = bodies strictly above the centre
    while b > 0:
        b = int(m.parent[b])
        path.append(b)
    assert sorted(rev) == sorted(path)
    assert np.all(m.armature[:6] == 0)


def test_host_struct_shapes():
    m = HumanoidModel()
    h = m.host_struct()
    assert h.nshape == 1 and h.nvert == len(m.hull)
    h2 = m.host_struct([m, HumanoidModel(scale=np.full(NB, 1.1))])
    assert h2.nshape == 2
