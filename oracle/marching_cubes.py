"""TEST INFRASTRUCTURE (oracle): numpy restatement of the marching cubes of hold_b200 (hold_mc_*; tables from
tools/gen_mc_tables.py -> oracle/mc_tables.json).  The reference calls skimage.measure.marching_cubes_lewiner
(utils/meshing.py:51); skimage is not in this image, so this is NOT a restatement of Lewiner's algorithm and the GPU kernels are
unpinned at that boundary: they are held to this restatement bit for bit and to geometric properties (closed, consistently
oriented, vertices on the level set, volume) in tests/test_cpu_mc.py / tests/test_gpu_mc.py.

Conventions: a grid node is INSIDE when value < level; a vertex sits on every grid edge whose ends differ, at lower_node +
(level - v0) / (v1 - v0) along the edge (float32), in INDEX coordinates; vertices are ordered by (node of the edge's lower end in C
order, axis), faces by (cell in C order, table order); normals (right-hand rule) point towards increasing values."""
import json
import os

import numpy as np

_T = None


def tables():
    global _T
    if _T is None:
        d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mc_tables.json")))
        _T = (np.asarray(d["ntri"], np.int32), np.asarray(d["tri"], np.int32), d["width"])
    return _T


def edge_origin_axis(e):
    """edge id -> (dx, dy, dz of its lower end inside the cell, axis)"""
    axis, r = e // 4, e % 4
    a, b = r & 1, r >> 1
    o = [0, 0, 0]
    u, v = [k for k in range(3) if k != axis]
    o[u], o[v] = a, b
    return o, axis


def marching_cubes(vol, level=0.0):
    vol = np.ascontiguousarray(vol, np.float32)
    level = np.float32(level)
    n0, n1, n2 = vol.shape
    ins = vol < level
    # vertex slots: [node, axis]
    slot = np.zeros((n0, n1, n2, 3), bool)
    slot[:-1, :, :, 0] = ins[:-1] != ins[1:]
    slot[:, :-1, :, 1] = ins[:, :-1] != ins[:, 1:]
    slot[:, :, :-1, 2] = ins[:, :, :-1] != ins[:, :, 1:]
    flat = slot.reshape(-1)
    vid = np.cumsum(flat, dtype=np.int64) - flat            # exclusive scan
    idx = np.nonzero(flat)[0]
    node, axis = idx // 3, idx % 3
    i, j, k = node // (n1 * n2), (node // n2) % n1, node % n2
    v0 = vol[i, j, k]
    v1 = vol[i + (axis == 0), j + (axis == 1), k + (axis == 2)]
    t = ((level - v0) / (v1 - v0)).astype(np.float32)
    verts = np.stack([i, j, k], 1).astype(np.float32)
    verts[np.arange(idx.size), axis] += t
    # cells
    case = np.zeros((n0 - 1, n1 - 1, n2 - 1), np.int32)
    for c in range(8):
        x, y, z = c & 1, (c >> 1) & 1, (c >> 2) & 1
        case |= ins[x:n0 - 1 + x, y:n1 - 1 + y, z:n2 - 1 + z].astype(np.int32) << c
    ntri, tri, width = tables()
    cflat = case.reshape(-1)
    cells = np.nonzero(ntri[cflat])[0]
    ci, cj, ck = cells // ((n1 - 1) * (n2 - 1)), (cells // (n2 - 1)) % (n1 - 1), cells % (n2 - 1)
    faces = []
    eo = [edge_origin_axis(e) for e in range(12)]
    for s in range(width):
        m = ntri[cflat[cells]] > s
        if not m.any():
            break
        f = np.zeros((int(m.sum()), 3), np.int64)
        for q in range(3):
            e = tri[cflat[cells[m]], 3 * s + q]
            ox = np.array([eo[x][0][0] for x in range(12)])[e]
            oy = np.array([eo[x][0][1] for x in range(12)])[e]
            oz = np.array([eo[x][0][2] for x in range(12)])[e]
            ax = np.array([eo[x][1] for x in range(12)])[e]
            nd = ((ci[m] + ox) * n1 + (cj[m] + oy)) * n2 + (ck[m] + oz)
            f[:, q] = vid[nd * 3 + ax]
        faces.append((cells[m], np.full(int(m.sum()), s), f))
    if not faces:
        return verts, np.zeros((0, 3), np.int32)
    cid = np.concatenate([a for a, _, _ in faces]); sid = np.concatenate([b for _, b, _ in faces]); ff = np.concatenate([c for _, _, c in faces])
    order = np.lexsort((sid, cid))
    return verts, ff[order].astype(np.int32)
