"""Generates the marching-cubes case tables of hold_b200/csrc/mc_tables.h (and the same tables for the oracle / tests as
oracle/mc_tables.json) instead of copying a published 256 x 16 table: every case is derived from first principles.

Cube corner i = (x, y, z) = (i & 1, (i >> 1) & 1, (i >> 2) & 1); edge id = 4 * axis + (a + 2 * b) with (a, b) the corner's other two
coordinates in increasing axis order.  A corner is INSIDE when its value is below the level (negative SDF).
Per face (corners counter-clockwise seen from outside the cube) the crossings of the iso-contour are joined by segments directed from
the crossing where the boundary walk LEAVES the inside to the crossing where it ENTERS it again; on an ambiguous face (two diagonal
inside corners) every inside corner is cut off on its own.  That rule depends only on the face's four corner states, so the two cubes
sharing a face draw the same segments: the surface is watertight by construction.  Adjacent faces walk their shared edge in opposite
directions, so segments chain into closed directed loops; each loop is fanned into triangles, reversed so that the normals
(right-hand rule) point to the OUTSIDE (increasing value)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def corner(i):
    return (i & 1, (i >> 1) & 1, (i >> 2) & 1)


def edge_id(c0, c1):
    p, q = corner(c0), corner(c1)
    axis = [k for k in range(3) if p[k] != q[k]]
    assert len(axis) == 1
    axis = axis[0]
    others = [p[k] for k in range(3) if k != axis]
    return 4 * axis + others[0] + 2 * others[1]


def faces():
    """six faces, each as 4 corner ids counter-clockwise seen from outside"""
    out = []
    for axis in range(3):
        for side in (0, 1):
            u, v = [k for k in range(3) if k != axis]
            quad = []
            for (a, b) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                c = [0, 0, 0]
                c[axis], c[u], c[v] = side, a, b
                quad.append(c[0] | (c[1] << 1) | (c[2] << 2))
            # orientation: (p1 - p0) x (p2 - p1) must equal the outward normal
            p = [corner(i) for i in quad]
            e1 = [p[1][k] - p[0][k] for k in range(3)]
            e2 = [p[2][k] - p[1][k] for k in range(3)]
            n = [e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]]
            want = -1 if side == 0 else 1
            if n[axis] != want:
                quad = quad[::-1]
            out.append(quad)
    return out


FACES = faces()


def case_triangles(case):
    inside = [(case >> i) & 1 for i in range(8)]
    nxt = {}
    for quad in FACES:
        exits, enters = [], []
        for k in range(4):
            a, b = quad[k], quad[(k + 1) % 4]
            if inside[a] and not inside[b]:
                exits.append((k, edge_id(a, b)))
            if not inside[a] and inside[b]:
                enters.append((k, edge_id(a, b)))
        # an inside run starts at an `enter` crossing and ends at the next `exit` crossing along the walk: join exit -> the enter that
        # opened ITS run (the last enter before it, cyclically).  On the ambiguous face this cuts each inside corner off on its own.
        for (ke, ee) in exits:
            best = max(enters, key=lambda t: ((t[0] - ke - 1) % 4))   # the enter closest BEFORE the exit in walk order
            assert ee not in nxt
            nxt[ee] = best[1]
    tris, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, e = [], start
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == start and len(loop) >= 3
        loop = loop[::-1]                       # outward normals
        tris += triangulate(loop)
    return tris


def edge_faces(e):
    """the two cube faces (axis, side) a cube edge lies on"""
    axis, r = e // 4, e % 4
    u, v = [k for k in range(3) if k != axis]
    return {(u, r & 1), (v, r >> 1)}


def triangulate(loop):
    """All triangulations of the loop polygon are enumerated; the one with the fewest diagonals that lie IN a cube face is taken
    (a diagonal between two vertices of the same face would put a triangle flat against that face, where the neighbouring cell may
    put the same one: a doubled, non-manifold edge), ties by the lexicographically smallest triangle list.  A plain fan produced 14
    such doubled edges on a 14 x 15 x 16 noise grid."""
    n = len(loop)
    best = None

    def rec(poly):
        if len(poly) == 3:
            yield [tuple(poly)]
            return
        a, b = poly[0], poly[1]                 # the triangle on edge (poly[0], poly[1]) picks its apex
        for k in range(2, len(poly)):
            left, right = poly[1:k + 1], [poly[0]] + poly[k:]
            for l in (rec(left) if len(left) >= 3 else [[]]):
                for r in (rec(right) if len(right) >= 3 else [[]]):
                    yield [(a, b, poly[k])] + l + r

    adjacent = {frozenset((loop[i], loop[(i + 1) % n])) for i in range(n)}
    for cand in rec(list(loop)):
        bad = 0
        for t in cand:
            for x, y in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                if frozenset((x, y)) not in adjacent and edge_faces(x) & edge_faces(y):
                    bad += 1
        key = (bad, cand)
        if best is None or key < best:
            best = key
    return best[1]


def tables():
    tri = [case_triangles(c) for c in range(256)]
    width = max(len(t) for t in tri)
    edges = [sum(1 << e for e in {e for t in tr for e in t}) for tr in tri]
    return tri, width, edges


def main():
    tri, width, edges = tables()
    ntri = [len(t) for t in tri]
    flat = []
    for t in tri:
        row = [e for tr in t for e in tr] + [-1] * (3 * (width - len(t)))
        flat.append(row)
    with open(os.path.join(ROOT, "oracle", "mc_tables.json"), "w") as f:
        json.dump({"width": width, "ntri": ntri, "tri": flat, "edge_mask": edges}, f)
    with open(os.path.join(ROOT, "hold_b200", "csrc", "mc_tables.h"), "w") as f:
        f.write("// GENERATED by tools/gen_mc_tables.py (derivation and conventions there) -- do not edit.\n#pragma once\n"
                "#if defined(__CUDACC__)\n#define HOLD_MC_CONST __device__ const\n#else\n#define HOLD_MC_CONST static const\n#endif\nnamespace hold {\n")
        f.write(f"constexpr int kMcMaxTri = {width};\n")
        f.write("HOLD_MC_CONST unsigned char kMcNTri[256] = {" + ", ".join(map(str, ntri)) + "};\n")
        f.write(f"HOLD_MC_CONST signed char kMcTri[256][{3 * width}] = {{\n")
        for row in flat:
            f.write("  {" + ", ".join(map(str, row)) + "},\n")
        f.write("};\n")
        eo = []
        for e in range(12):
            axis, r = e // 4, e % 4
            o = [0, 0, 0]
            u, v = [k for k in range(3) if k != axis]
            o[u], o[v] = r & 1, r >> 1
            eo.append(o + [axis])
        f.write("// cube edge -> (offset of its lower end inside the cell, axis)\n")
        f.write("HOLD_MC_CONST unsigned char kMcEdge[12][4] = {" + ", ".join("{" + ", ".join(map(str, q)) + "}" for q in eo) + "};\n")
        f.write("}  // namespace hold\n")
    print("max triangles per cell", width, "total triangles over the 256 cases", sum(ntri))


if __name__ == "__main__":
    main()
