"""Generates the committed golden fixtures under tests/golden/ (run from the repo root, in the build
container, where /root/reference is mounted):

  config1_30x20_phs3.npz   README Laplace problem: cloud arrays, a sample of K, q, the solution of
                           the reference formulation (inv + GEMM + QR on the CPU oracle)
  config2_35x35_periodic.npz  periodic adv-diff cloud (demos/Advection/01): cloud arrays, K sample
  mesh_msh_cloud_{vel,phi,alln}.npz  the reference's own fixture updes/tests/data/mesh.msh parsed with the
                           oracle's literal restatement of GmshCloud (sorted nodes, normals, counts,
                           facet nodes) for the two facet-type sets of demos/NavierStokes/30_...:40-41

These vectors come from the ORACLE (round 1, when the reference could not be run here).  The files made by the reference's
own code are tests/golden/ref_*.npz (make_reference_golden.py); tests/test_reference_golden.py checks that both agree
(cloud arrays bit for bit, the config-1 solution to 8e-10).  The mesh fixture is the one piece of reference *data* on
this path and is stored in parsed form only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(0)


def cloud_arrays(c):
    names = list(c.facet_nodes.keys())
    return dict(sorted_nodes=c.sorted_nodes, sorted_outward_normals=c.sorted_outward_normals,
                counts=np.array([c.N, c.Ni, c.Nd, c.Nn, c.Nr]), Np=np.array(c.Np, dtype=np.int64),
                facet_names=np.array(names), facet_sizes=np.array([len(c.facet_nodes[k]) for k in names]),
                facet_nodes=np.concatenate([np.asarray(c.facet_nodes[k], dtype=np.int64) for k in names]),
                facet_types=np.array([c.facet_types[k] for k in names]),
                # original (grid / mesh) id of every sorted node: renumbering_map is old -> new (cloud.py:165)
                old_of_new=np.array([o for o, _ in sorted(c.renumbering_map.items(), key=lambda kv: kv[1])], dtype=np.int64))


c1 = O.RefSquareCloud(30, 20, {"South": "n", "West": "d", "North": "d", "East": "d"})
coef = np.tile([0.0, 0, 0, 1.0, 1.0], (c1.Ni, 1))
K = O.assemble_K(c1, "polyharmonic", 1, 3, coef)
rows = np.sort(rng.choice(c1.N + 3, 64, replace=False)); cols = np.sort(rng.choice(c1.N + 3, 64, replace=False))
xy = c1.sorted_nodes
bc = {f: (np.sin(np.pi * xy[ids, 0]) if f == "North" else np.zeros(len(ids))) for f, ids in c1.facet_nodes.items()}
q = O.assemble_q(c1, np.zeros(c1.Ni), bc)
vals, coeffs, _ = O.reference_solve(c1, "polyharmonic", 1, 1, coef, q)
np.savez_compressed(os.path.join(OUT, "config1_30x20_phs3.npz"), rows=rows, cols=cols, K_sample=K[rows][:, cols], q=q,
                    vals=vals, coeffs=coeffs, **cloud_arrays(c1))

c2 = O.RefSquareCloud(35, 35, {"South": "p1", "North": "p1", "West": "p2", "East": "p2"}, noise_seed=7)
coef2 = np.tile([1e4, 100.0, 0.0, -0.08, -0.08], (c2.Ni, 1))
K2 = O.assemble_K(c2, "polyharmonic", 1, 1, coef2)
rows = np.sort(np.concatenate([rng.choice(c2.Ni, 32, replace=False), np.arange(c2.Ni, c2.N + 1)]))
cols = np.sort(rng.choice(c2.N + 1, 96, replace=False))
np.savez_compressed(os.path.join(OUT, "config2_35x35_periodic.npz"), rows=rows, cols=cols, K_sample=K2[rows][:, cols],
                    **cloud_arrays(c2))

mesh = "/root/reference/updes/tests/data/mesh.msh"
for tag, ft in (("vel", {"Wall": "d", "Inflow": "d", "Outflow": "n", "Blowing": "d", "Suction": "d"}),
                ("phi", {"Wall": "n", "Inflow": "n", "Outflow": "d", "Blowing": "n", "Suction": "n"}),
                # updes/tests/test_operators.py:26 -- every facet Neumann
                ("alln", {"Wall": "n", "Inflow": "n", "Outflow": "n", "Blowing": "n", "Suction": "n"})):
    c = O.RefGmshCloud(mesh, ft)
    np.savez_compressed(os.path.join(OUT, "mesh_msh_cloud_%s.npz" % tag), **cloud_arrays(c))
print("golden fixtures written to", OUT)
