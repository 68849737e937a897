"""Writes a small Gmsh-4.0 ASCII mesh of a rectangular channel [0,1.5]x[0,1] (own data, used to test
the .msh reader without the reference's fixture).  Physical curves: Wall (bottom+top), Inflow (left),
Outflow (right); surface: Fluid.  Layout follows the Gmsh 4.0 format the reference parses
(updes/cloud.py:578-694): $PhysicalNames, $Entities, $Nodes, $Elements."""
import numpy as np


def write_channel_msh(path, nx=13, ny=9, lx=1.5, ly=1.0):
    xs, ys = np.linspace(0, lx, nx), np.linspace(0, ly, ny)
    tag = {}
    nodes = []          # (entity_dim, entity_tag, x, y)

    def add(dim, ent, x, y):
        nodes.append((dim, ent, x, y)); tag[(round(x, 12), round(y, 12))] = len(nodes)

    corners = [(0, 0), (lx, 0), (lx, ly), (0, ly)]
    for k, (x, y) in enumerate(corners):
        add(0, k + 1, x, y)
    for x in xs[1:-1]: add(1, 1, x, 0.0)        # bottom wall
    for y in ys[1:-1]: add(1, 2, lx, y)         # outflow
    for x in xs[1:-1][::-1]: add(1, 3, x, ly)   # top wall
    for y in ys[1:-1][::-1]: add(1, 4, 0.0, y)  # inflow
    for x in xs[1:-1]:
        for y in ys[1:-1]: add(2, 1, x, y)
    phys = {1: "Wall", 2: "Inflow", 3: "Outflow", 4: "Fluid"}
    curve_phys = {1: 1, 2: 3, 3: 1, 4: 2}
    curve_pts = {1: (1, 2), 2: (2, 3), 3: (3, 4), 4: (4, 1)}
    with open(path, "w") as f:
        f.write("$MeshFormat\n4 0 8\n$EndMeshFormat\n$PhysicalNames\n%d\n" % len(phys))
        for k, name in phys.items():
            f.write('%d %d "%s"\n' % (2 if name == "Fluid" else 1, k, name))
        f.write("$EndPhysicalNames\n$Entities\n4 4 1 0\n")
        for k, (x, y) in enumerate(corners):
            f.write("%d %g %g 0 %g %g 0 0\n" % (k + 1, x, y, x, y))
        for c in range(1, 5):
            a, b = curve_pts[c]
            f.write("%d 0 0 0 %g %g 0 1 %d 2 %d %d\n" % (c, lx, ly, curve_phys[c], a, -b))
        f.write("1 0 0 0 %g %g 0 1 4 4 1 2 3 4\n$EndEntities\n" % (lx, ly))
        blocks = {}
        for i, (dim, ent, x, y) in enumerate(nodes):
            blocks.setdefault((dim, ent), []).append((i + 1, x, y))
        f.write("$Nodes\n%d %d\n" % (len(blocks), len(nodes)))
        for (dim, ent), items in blocks.items():
            f.write("%d %d 0 %d\n" % (ent, dim, len(items)))
            for (i, x, y) in items:
                f.write("%d %.16g %.16g 0\n" % (i, x, y))
        f.write("$EndNodes\n")
        # line elements along each curve (needed to attach the corner points to facets)
        def chain(pts):
            return [tag[(round(x, 12), round(y, 12))] for x, y in pts]
        curves = {1: chain([(x, 0.0) for x in xs]), 2: chain([(lx, y) for y in ys]),
                  3: chain([(x, ly) for x in xs[::-1]]), 4: chain([(0.0, y) for y in ys[::-1]])}
        nel = sum(len(v) - 1 for v in curves.values())
        f.write("$Elements\n4 %d\n" % nel)
        eid = 1
        for c, ch in curves.items():
            f.write("%d 1 1 %d\n" % (c, len(ch) - 1))
            for a, b in zip(ch[:-1], ch[1:]):
                f.write("%d %d %d\n" % (eid, a, b)); eid += 1
        f.write("$EndElements\n")


if __name__ == "__main__":
    import sys
    write_channel_msh(sys.argv[1] if len(sys.argv) > 1 else "channel.msh")
