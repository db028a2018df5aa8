"""Marching cubes, CPU side: the triangle table against the reference's copy (when the tree is present) and the oracle's
extraction on an analytic sphere (closed surface, every vertex on a grid edge at the iso level)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/MarchingCubes"


def product_table():
    text = open(os.path.join(ROOT, "tsdf_b200", "csrc", "mc_tables.h")).read()
    rows = re.findall(r'"([0-9a-b]*)"', text.split("kMcTriangles[256] = {")[1].split("};")[0])
    assert len(rows) == 256
    return rows


def test_table_shape():
    rows = product_table()
    assert rows[0] == "" and rows[255] == ""
    assert all(len(r) % 3 == 0 and len(r) <= 15 for r in rows)
    # complementary cube types cut the same edges
    assert all(sorted(set(rows[t])) == sorted(set(rows[255 - t])) for t in range(256))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "MC_triangle_table.cu")), reason="reference tree not present")
def test_table_equals_reference():
    src = open(os.path.join(REF, "MC_triangle_table.cu")).read()
    ref = []
    for r in re.findall(r"\{([-\d,\s]+)\}", src[src.index("TRIANGLE_TABLE"):]):
        v = [int(x) for x in r.split(",") if x.strip()]
        if len(v) == 16:
            ref.append("".join("%x" % e for e in v if e >= 0))
    assert len(ref) == 256 and ref == product_table()
    counts = [int(x) for x in re.search(r"VERTICES_FOR_CUBE_TYPE\[256\]\s*=\s*\{([^}]*)\}", src).group(1).split(",") if x.strip()]
    assert counts == [len(r) for r in product_table()]
    edges = open(os.path.join(REF, "MC_edge_table.cu")).read()
    pairs = re.findall(r"\{\s*(\d+),\s*(\d+)\s*\}", edges[edges.index("EDGE_VERTICES"):])[:12]
    mine = re.findall(r"\{(\d+), (\d+)\}", open(os.path.join(ROOT, "tsdf_b200", "csrc", "mc_tables.h")).read().split("kMcEdgeCorners[12][2]")[1])
    assert [(int(a), int(b)) for a, b in pairs] == [(int(a), int(b)) for a, b in mine]


def test_oracle_sphere(built):
    from oracle import oracle
    n = (24, 20, 28)
    vox = np.array([10.0, 12.0, 9.0], np.float32)
    off = np.array([5.0, -3.0, 100.0], np.float32)
    z, y, x = np.meshgrid(*(np.arange(m) + 0.5 for m in (n[2], n[1], n[0])), indexing="ij")
    centre = np.array([120.0, 120.0, 126.0])
    d = np.sqrt((x * vox[0] - centre[0]) ** 2 + (y * vox[1] - centre[1]) ** 2 + (z * vox[2] - centre[2]) ** 2) - 70.0
    v = oracle.mc_extract(d.astype(np.float32).reshape(-1), n, vox, off)
    assert len(v) > 0 and len(v) % 3 == 0
    r = np.linalg.norm(v - off - centre, axis=1)
    assert np.all(np.abs(r - 70.0) < 3.0)                  # linear interpolation error on a 10 mm grid
    # closed surface: every undirected edge is shared by exactly two triangles
    tri = v.reshape(-1, 3, 3)
    keys = {}
    for t in tri:
        p = [tuple(np.round(q, 3)) for q in t]
        for a, b in ((0, 1), (1, 2), (2, 0)):
            if p[a] == p[b]:
                continue
            k = tuple(sorted((p[a], p[b])))
            keys[k] = keys.get(k, 0) + 1
    assert all(c == 2 for c in keys.values())
