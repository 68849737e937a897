"""Clouds of collocation nodes: the input contract of the hot path.

Re-statement of ``updes/cloud.py`` (reference) for the *global* path (``support_size="max"``), in
vectorised numpy: O(N) work and memory, no ``(N, N-1)`` support table (reference cloud.py:108-112
materialises one, 64.8 GB at N = 90 000).  What the assembly consumes is reproduced exactly:

* ``sorted_nodes`` (N, 2), renumbered internal -> dirichlet -> neumann -> robin -> periodic classes
  sorted by suffixed type, ascending original id inside each class      (cloud.py:115-172)
* ``sorted_outward_normals`` for n / r / p nodes, indexed ``[i - Ni - Nd]`` (cloud.py:407-408)
* counts ``N, Ni, Nd, Nn, Nr, Np`` (list per periodic group), ``facet_nodes``, ``facet_types`` with the
  facet index appended to periodic types (cloud.py:43-49), ``renumbering_map``.

``SquareCloud``  : cloud.py:378-510.     ``GmshCloud`` : cloud.py:531-734 (Gmsh 4.0 ASCII reader).
Every node's support is "all other nodes" (cloud.py:110-112 drops the node itself), which the
assembly encodes as a skipped diagonal entry rather than an index table.
"""
from __future__ import annotations

import os

import numpy as np


class _LazySupports:
    """Read-only mapping node id -> list of the other nodes by increasing distance (ties by id), computed per access."""

    def __init__(self, nodes, support_size=None):
        self._nodes = nodes
        self._keep = (nodes.shape[0] if support_size is None else int(support_size)) - 1      # cloud.py:110-112: k nearest incl. self, self dropped

    def __len__(self):
        return self._nodes.shape[0]

    def __iter__(self):
        return iter(range(len(self)))

    def keys(self):
        return range(len(self))

    def __contains__(self, i):
        return isinstance(i, (int, np.integer)) and 0 <= i < len(self)

    def __getitem__(self, i):
        n = len(self)
        i = int(i)
        if not 0 <= i < n:
            raise KeyError(i)
        d = self._nodes - self._nodes[i]
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
        order = np.lexsort((np.arange(n), d2))
        return [int(j) for j in order if j != i][: self._keep]

    def items(self):
        return ((i, self[i]) for i in range(len(self)))

    def values(self):
        return (self[i] for i in range(len(self)))


class Cloud:
    """Base bookkeeping shared by all clouds (reference cloud.py:10-50)."""

    def __init__(self, facet_types, support_size="max"):
        # "max", None, or the number of nodes N (what the reference turns "max" into, cloud.py:97-98) select the global
        # path; anything smaller drops the farthest nodes from every support (RBF-FD): checked once N is known
        if not (support_size in ("max", None) or (isinstance(support_size, (int, np.integer)) and not isinstance(support_size, bool))):
            raise ValueError("support_size must be 'max', None or an integer, got %r" % (support_size,))
        self._requested_support = support_size
        self.support_size = "max"
        self.N = self.Ni = self.Nd = self.Nr = self.Nn = 0
        self.Np = []
        self.dim = 2
        self.facet_precedence = {k: i for i, (k, v) in enumerate(facet_types.items())}
        self.facet_types = {k: (v + str(i) if v[0] == "p" else v) for i, (k, v) in enumerate(facet_types.items())}
        self.facet_nodes = {}
        self.node_types = {}
        self.outward_normals = {}
        self.renumbering_map = {}

    # -- renumbering (cloud.py:115-172) -----------------------------------------------------------
    def _renumber(self, types, coords, normals_by_old, facet_nodes_old):
        """types: array of str per original id; returns the sorted arrays and fills the dicts."""
        req = getattr(self, "_requested_support", "max")
        N = len(types)
        # cloud.py:97-100: "max" means all N nodes.  A smaller support (RBF-FD) does not change the geometry built here,
        # so the cloud itself is constructed (the reference's own test_interpolation.py asks for N - 1 and only permutes
        # fields); assembling on it is refused (assembly.DeviceRows): the product implements the global path only.
        self.support_size = N if req in ("max", None) else int(req)
        if not 0 < self.support_size <= max(N, 1):
            raise AssertionError("Support size must be strictly greater than 0 and at most the number of nodes")
        first = np.array([t[0] for t in types])
        order = [np.flatnonzero(first == c) for c in ("i", "d", "n", "r")]
        p_ids = np.flatnonzero(first == "p")
        for key in sorted(set(types[p_ids])):
            order.append(np.flatnonzero(types == key))
        old_of_new = np.concatenate(order) if N else np.zeros(0, dtype=int)
        if len(old_of_new) != N:
            raise ValueError("Unknown node type")
        new_of_old = np.empty(N, dtype=np.int64)
        new_of_old[old_of_new] = np.arange(N)
        self._new_of_old = new_of_old
        self._old_of_new = old_of_new
        self.renumbering_map = dict(zip(range(N), new_of_old.tolist()))
        self.sorted_nodes = np.ascontiguousarray(coords[old_of_new], dtype=np.float64)
        sorted_types = types[old_of_new]
        self.node_types = dict(zip(range(N), sorted_types.tolist()))
        self.facet_nodes = {f: new_of_old[np.asarray(ids, dtype=np.int64)].tolist() for f, ids in facet_nodes_old.items()}
        # normals exist for n / r / p nodes, which are contiguous after the Dirichlet block
        nb = self.Nn + self.Nr + sum(self.Np)
        start = self.Ni + self.Nd
        if nb > 0:
            son = np.zeros((nb, 2))
            for old, nv in normals_by_old.items():
                son[new_of_old[old] - start] = nv
            self.sorted_outward_normals = son
            self.outward_normals = {start + k: son[k] for k in range(nb)}
        else:
            self.sorted_outward_normals = np.zeros((0, 2))
            self.outward_normals = {}

    @classmethod
    def from_arrays(cls, sorted_nodes, counts, Np, facet_types, facet_nodes, sorted_outward_normals, old_of_new=None):
        """Cloud from already-renumbered arrays (user-supplied point sets, or clouds built elsewhere).

        ``counts`` = (N, Ni, Nd, Nn, Nr); ``facet_types`` must already carry the reference's suffix on
        periodic types; ``facet_nodes`` maps facet -> sorted node ids; normals are indexed
        ``[i - Ni - Nd]`` as in the reference (cloud.py:407-408)."""
        self = cls.__new__(cls)
        Cloud.__init__(self, {}, support_size="max")
        self.N, self.Ni, self.Nd, self.Nn, self.Nr = (int(c) for c in counts)
        self.Np = [int(v) for v in Np]
        self.facet_types = dict(facet_types)
        self.facet_precedence = {k: i for i, k in enumerate(self.facet_types)}
        self.facet_nodes = {k: [int(i) for i in v] for k, v in facet_nodes.items()}
        self.sorted_nodes = np.ascontiguousarray(sorted_nodes, dtype=np.float64)
        self.sorted_outward_normals = np.ascontiguousarray(sorted_outward_normals, dtype=np.float64).reshape(-1, 2)
        start = self.Ni + self.Nd
        self.outward_normals = {start + k: self.sorted_outward_normals[k] for k in range(len(self.sorted_outward_normals))}
        types = ["i"] * self.Ni + ["d"] * self.Nd + ["n"] * self.Nn + ["r"] * self.Nr
        types += ["p"] * (self.N - len(types))
        for f, ids in self.facet_nodes.items():
            for i in ids:
                types[i] = self.facet_types[f]
        self.node_types = dict(enumerate(types))
        # old_of_new[i] = original (mesh / grid) id of sorted node i; lets interpolate_field relate two clouds
        # built on the same nodes with different boundary types
        oon = np.arange(self.N) if old_of_new is None else np.asarray(old_of_new, dtype=np.int64)
        self._old_of_new = oon
        self._new_of_old = np.empty(self.N, dtype=np.int64)
        self._new_of_old[oon] = np.arange(self.N)
        self.renumbering_map = {int(o): int(k) for k, o in enumerate(oon)}     # old -> new, in new-id order (cloud.py:165)
        return self

    @property
    def nodes(self):
        """dict new id -> coordinates (reference attribute ``Cloud.nodes`` after renumbering)."""
        return {i: self.sorted_nodes[i] for i in range(self.N)}

    @property
    def local_supports(self):
        """``cloud.local_supports[i]``: the other N - 1 nodes ordered by distance from node i (reference cloud.py:83-112 with
        ``support_size="max"``; demos read it to pick a node's neighbourhood, e.g. demos/Advection/00_...:68).  The reference
        stores all N lists; here a list is computed when it is asked for (O(N log N) each), so a 250k-node cloud does not
        carry an (N, N - 1) table.  Equidistant nodes are ordered by node id (the reference takes BallTree's order there,
        which is implementation-defined)."""
        ss = getattr(self, "support_size", self.N)
        return _LazySupports(self.sorted_nodes, self.N if ss in ("max", None) else int(ss))

    @property
    def sorted_local_supports(self):
        """(N, N - 1) int array of the lists above (cloud.py:403-408).  Small clouds only."""
        if self.N > 20000:
            raise MemoryError("sorted_local_supports is an (N, N-1) table; the global path never needs it (N = %d)" % self.N)
        ls = self.local_supports
        return np.array([ls[i] for i in range(self.N)], dtype=np.int64).reshape(self.N, -1) if self.N > 1 else np.zeros((self.N, 0), dtype=np.int64)

    def sort_dict_by_keys(self, dictionary):
        """cloud.py:72-81"""
        items = sorted(dictionary.items(), key=lambda kv: kv[0])
        return np.stack([np.asarray(v) for _, v in items], axis=0)

    # ---- plotting helpers of the reference surface (cloud.py:175-375): host-only, outside the hot path -------------
    @staticmethod
    def _pyplot():
        try:
            import matplotlib.pyplot as plt
        except ImportError as e:
            raise ImportError("Cloud.visualize_* needs matplotlib, which is not installed; the solver itself does not") from e
        return plt

    def average_spacing(self):
        """Mean distance over all node pairs, the node itself included (cloud.py:62-70)."""
        xy = np.asarray(self.sorted_nodes)
        d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
        iu = np.triu_indices(self.N)
        return float(d[iu].mean())

    def visualize_cloud(self, ax=None, title="Cloud", xlabel=r"$x$", ylabel=r"$y$", legend_size=8, figsize=(5.5, 5), **kwargs):
        """Scatter plot of the nodes coloured by type (cloud.py:175-206)."""
        plt = self._pyplot()
        if ax is None:
            ax = plt.figure(figsize=figsize).add_subplot(1, 1, 1)
        xy = np.asarray(self.sorted_nodes)
        Ni, Nd, Nn = self.Ni, self.Nd, self.Nn
        for lo, hi, colour, label in ((0, Ni, "w", "internal"), (Ni, Ni + Nd, "r", "dirichlet"), (Ni + Nd, Ni + Nd + Nn, "g", "neumann"),
                                      (Ni + Nd + Nn, self.N, "b", "robin")):
            if hi > lo:
                ax.scatter(x=xy[lo:hi, 0], y=xy[lo:hi, 1], c=colour, label=label, **kwargs)
        if xlabel:
            ax.set_xlabel(xlabel)
        if ylabel:
            ax.set_ylabel(ylabel)
        ax.set_title(title)
        ax.legend(bbox_to_anchor=(1.0, 0.5), loc="center left", prop={"size": legend_size})
        plt.tight_layout()
        return ax

    def visualize_normals(self, ax=None, title="Normal vectors", xlabel=r"$x$", ylabel=r"$y$", figsize=(5.5, 5), zoom_region=None, **kwargs):
        """Outward normals on Neumann / Robin / periodic nodes (cloud.py:209-238)."""
        plt = self._pyplot()
        if ax is None:
            ax = plt.figure(figsize=figsize).add_subplot(1, 1, 1)
        normals = np.asarray(self.sorted_outward_normals, dtype=np.float64).reshape(-1, 2)
        if normals.shape[0] == 0:
            return ax
        first = self.Ni + self.Nd
        xy = np.asarray(self.sorted_nodes)[first:first + normals.shape[0]]
        q = ax.quiver(xy[:, 0], xy[:, 1], normals[:, 0] / 100, normals[:, 1] / 100, color="w", label="normals", **kwargs)
        ax.quiverkey(q, X=0.5, Y=1.1, U=1, label="Normals", labelpos="E")
        ax.scatter(x=xy[:, 0], y=xy[:, 1], c="m", **kwargs)
        if xlabel:
            ax.set_xlabel(xlabel)
        if ylabel:
            ax.set_ylabel(ylabel)
        ax.set_title(title)
        if zoom_region:
            ax.set_xlim((zoom_region[0], zoom_region[1]))
            ax.set_ylim((zoom_region[2], zoom_region[3]))
        plt.tight_layout()
        return ax

    def visualize_field(self, field, projection="2d", title="Field", xlabel=r"$x$", ylabel=r"$y$", levels=50, colorbar=True, ax=None,
                        figsize=(6, 5), extend="neither", **kwargs):
        """Filled contours (2d) or a triangulated surface (3d) of a nodal field (cloud.py:241-285); returns (ax, image)."""
        plt = self._pyplot()
        xy = np.asarray(self.sorted_nodes)
        field = np.asarray(field, dtype=np.float64)
        if field.ndim > 1:
            field = field[:, 0, ...]
        if ax is None:
            fig = plt.figure(figsize=figsize)
            ax = fig.add_subplot(1, 1, 1, projection="3d") if projection == "3d" else fig.add_subplot(1, 1, 1)
        if projection == "3d":
            img = ax.plot_trisurf(xy[:, 0], xy[:, 1], field, **kwargs)
        else:
            img = ax.tricontourf(xy[:, 0], xy[:, 1], field, levels=levels, extend=extend, **kwargs)
        if colorbar and projection != "3d":
            plt.sca(ax)
            plt.colorbar(img, extend=extend)
        if xlabel:
            ax.set_xlabel(xlabel)
        if ylabel:
            ax.set_ylabel(ylabel)
        ax.set_title(title)
        plt.tight_layout()
        return ax, img

    def animate_fields(self, fields, filename=None, titles="Field", xlabel=r"$x$", ylabel=r"$y$", levels=50, figsize=(6, 5),
                       cmaps="jet", cbarsplit=50, duration=5, colorbar=True, vmin=None, vmax=None, **kwargs):
        """One filled-contour animation per field history, stacked vertically (reference surface, cloud.py:276-358; the
        demos end with it).  ``fields``: list of histories, each a list of nodal fields or a (steps, N) array.  Saved with
        pillow (.gif) or ffmpeg (.mp4) when ``filename`` is given; returns the axes.  Host-only, needs matplotlib."""
        plt = self._pyplot()
        from matplotlib.animation import FuncAnimation
        histories = [np.stack([np.asarray(f, dtype=np.float64) for f in h], axis=0) if isinstance(h, (list, tuple))
                     else np.asarray(h, dtype=np.float64) for h in fields]
        x, y = np.asarray(self.sorted_nodes[:, 0]), np.asarray(self.sorted_nodes[:, 1])
        fig, axes = plt.subplots(len(histories), 1, figsize=figsize, sharex=True)
        axes = [axes] if len(histories) == 1 else list(axes)
        cmaps = list(cmaps) if isinstance(cmaps, (list, tuple)) else [cmaps] * len(histories)
        ranges = []
        for k, (h, ax) in enumerate(zip(histories, axes)):
            lo = float(np.min(h)) if vmin is None else vmin
            hi = float(np.max(h)) if vmax is None else vmax
            if hi <= lo:
                hi = lo + 1e-3                                   # a constant history still needs a colour range
            ranges.append((lo, hi))
            ax.tricontourf(x, y, h[0], levels=levels, vmin=lo, vmax=hi, cmap=cmaps[k], **kwargs)
            if colorbar:
                mappable = plt.cm.ScalarMappable(cmap=cmaps[k])
                mappable.set_array(h)
                mappable.set_clim(lo, hi)
                plt.colorbar(mappable, boundaries=np.linspace(lo, hi, cbarsplit), shrink=1.0, aspect=10, ax=ax)
            ax.set_title(titles[k] if isinstance(titles, (list, tuple)) and k < len(titles) else
                         (titles if isinstance(titles, str) and len(histories) == 1 else "field # %d" % (k + 1)))
            if k == len(histories) - 1:
                ax.set_xlabel(xlabel)
            ax.set_ylabel(ylabel)

        def draw(frame):
            return [ax.tricontourf(x, y, h[frame], levels=levels, vmin=r[0], vmax=r[1], cmap=c, extend="min", **kwargs)
                    for ax, h, r, c in zip(axes, histories, ranges, cmaps)]

        steps = histories[0].shape[0]
        anim = FuncAnimation(fig, draw, frames=steps, repeat=False, interval=100)
        plt.tight_layout()
        if filename:
            anim.save(filename, writer="ffmpeg" if str(filename).endswith(".mp4") else "pillow", fps=steps / duration)
            print("Animation saved at:", filename)
        return axes


class SquareCloud(Cloud):
    """Regular or jittered grid on the unit square (reference cloud.py:378-510).

    ``noise_key``: ``None`` for the regular grid, else an integer seed (or array whose entries are
    folded into one).  The reference draws the jitter from ``jax.random``; that stream cannot be
    reproduced without JAX, so jittered clouds use ``numpy.random.default_rng(seed)`` -- the rule is
    the same: uniform in +-min(dx, dy)/2 on every node that is not d / n / r (cloud.py:431-444).
    """

    def __init__(self, Nx=7, Ny=5, noise_key=None, **kwargs):
        super().__init__(**kwargs)
        for k in self.facet_types:
            if k not in ("North", "South", "East", "West"):
                raise KeyError("SquareCloud facets must be named North, South, East, West (cloud.py:455-468)")
        self.Nx, self.Ny, self.N = Nx, Ny, Nx * Ny
        gid = np.arange(self.N)
        I, J = gid // Ny, gid % Ny                       # cloud.py:411-422: gid = i*Ny + j
        self.global_indices = gid.reshape(Nx, Ny)
        self.global_indices_rev = None

        # node types with the reference's facet precedence N, S, E, W (cloud.py:455-468)
        facet_of = np.full(self.N, "", dtype=object)
        facet_of[I == 0] = "West"
        facet_of[I == Nx - 1] = "East"
        facet_of[J == 0] = "South"
        facet_of[J == Ny - 1] = "North"
        types = np.full(self.N, "i", dtype=object)
        facet_nodes_old = {k: [] for k in self.facet_types}
        for f in ("West", "East", "South", "North"):
            ids = np.flatnonzero(facet_of == f)
            if len(ids):
                types[ids] = self.facet_types[f]
                facet_nodes_old[f] = ids
        # counts (cloud.py:470-488)
        Np = {v[:-1]: 0 for v in self.facet_types.values() if v[0] == "p"}
        for f, t in self.facet_types.items():
            cnt = len(facet_nodes_old[f])
            if t == "d":
                self.Nd += cnt
            elif t == "n":
                self.Nn += cnt
            elif t == "r":
                self.Nr += cnt
            elif t[0] == "p":
                Np[t[:-1]] += cnt
        self.Np = [Np[k] for k in sorted(Np)]
        self.Ni = self.N - self.Nd - self.Nn - self.Nr - sum(self.Np)

        # coordinates (cloud.py:425-446)
        x = np.linspace(0, 1.0, Nx)
        y = np.linspace(0, 1.0, Ny)
        coords = np.stack([x[I], y[J]], axis=1)
        if noise_key is not None:
            seed = int(np.asarray(noise_key).astype(np.int64).ravel().sum()) if not isinstance(noise_key, (int, np.integer)) else int(noise_key)
            delta = min(x[1] - x[0], y[1] - y[0]) / 2.0
            noise = np.random.default_rng(seed).uniform(-delta, delta, size=(self.N, 2))
            movable = ~np.isin(types, ["d", "n", "r"])
            coords = coords + noise * movable[:, None]

        # outward normals of n / r / p nodes, same precedence (cloud.py:491-510)
        normals = {}
        first = np.array([t[0] for t in types])
        table = {"North": (0.0, 1.0), "South": (0.0, -1.0), "East": (1.0, 0.0), "West": (-1.0, 0.0)}
        for f, nv in table.items():
            for old in np.flatnonzero((facet_of == f) & np.isin(first, ["n", "r", "p"])):
                normals[int(old)] = np.array(nv)
        self._renumber(types, coords, normals, facet_nodes_old)
        # the reference renumbers its index grid too (cloud.py:153-157): global_indices[k, l] = the NEW id of grid node (k, l)
        self.global_indices = self._new_of_old[gid].reshape(Nx, Ny)
        self.global_indices_rev = {int(self.global_indices[k, l]): (k, l) for k in range(Nx) for l in range(Ny)} if self.N <= 20000 else None

    def print_global_indices(self):
        """cloud.py:51-58: the node ids laid out as the grid is drawn (y upwards)."""
        print(np.flip(self.global_indices.T, axis=0))


class GmshCloud(Cloud):
    """Cloud read from a Gmsh 4.0 ASCII ``.msh`` file (reference cloud.py:531-734).

    Only ``.msh`` input is supported: generating a mesh from a ``.py`` script needs the ``gmsh``
    package (cloud.py:573-575), which is outside the hot path.  Periodic facets are not available
    on Gmsh clouds in the reference either (``Np`` is never set, cloud.py:690-694).
    """

    def __init__(self, filename, mesh_save_location=None, **kwargs):
        super().__init__(**kwargs)
        if not str(filename).endswith(".msh"):
            raise NotImplementedError("GmshCloud reads .msh files; run gmsh yourself to produce one")
        self.filename = filename
        types, coords, facet_nodes_old, tag_nodes = self._read_msh()
        normals = self._normals(types, coords, tag_nodes)
        self._renumber(types, coords, normals, facet_nodes_old)
        self.facet_tag_nodes = {t: self._new_of_old[np.asarray(ids, dtype=np.int64)].tolist() for t, ids in tag_nodes.items()}

    @staticmethod
    def _section(lines, name):
        start = next(i for i, l in enumerate(lines) if l.strip() == "$" + name)
        end = next(i for i in range(start, len(lines)) if lines[i].strip() == "$End" + name)
        return lines[start + 1:end]

    def _read_msh(self):
        with open(self.filename, "r") as f:
            lines = f.read().splitlines()
        # physical names: all but the last one are facets (cloud.py:586-593)
        sec = self._section(lines, "PhysicalNames")
        phys = {}
        for l in sec[1:int(sec[0].split()[0])]:
            t = l.split()
            phys[int(t[1])] = t[2][1:-1]
        # curve entity -> physical name (cloud.py:596-605)
        sec = self._section(lines, "Entities")
        h = sec[0].split()
        n_points, n_curves = int(h[0]), int(h[1])
        self.facet_names = {}
        for l in sec[1 + n_points:1 + n_points + n_curves]:
            t = l.split()
            self.facet_names[int(t[0])] = phys[int(t[-4])]
        # nodes (cloud.py:608-648)
        sec = self._section(lines, "Nodes")
        self.N = int(sec[0].split()[1])
        coords = np.zeros((self.N, 2))
        types = np.full(self.N, "", dtype=object)
        facet_nodes_old = {v: [] for v in self.facet_names.values()}
        tag_nodes = {k: [] for k in self.facet_names}
        corners = {}
        k = 1
        while k < len(sec):
            t = sec[k].split()
            ent, dim, nb = int(t[0]), int(t[1]), int(t[-1])
            block = []
            for l in sec[k + 1:k + 1 + nb]:
                u = l.split()
                nid = int(u[0]) - 1
                coords[nid] = (float(u[1]), float(u[2]))
                if dim == 0:
                    corners[nid] = []
                elif dim == 1:
                    types[nid] = self.facet_types[self.facet_names[ent]]
                    block.append(nid)
                elif dim == 2:
                    types[nid] = "i"
            if dim == 1:
                facet_nodes_old[self.facet_names[ent]] += block
                tag_nodes[ent] += block
            k += 1 + nb
        # line elements decide which facets a corner touches (cloud.py:651-674)
        sec = self._section(lines, "Elements")
        k = 1
        while k < len(sec):
            t = sec[k].split()
            ent, dim, nb = int(t[0]), int(t[1]), int(t[-1])
            if dim == 1:
                for l in sec[k + 1:k + 1 + nb]:
                    ids = [int(v) - 1 for v in l.split()[1:]]
                    for c in corners:
                        if c in ids and any(v != c for v in ids):
                            corners[c].append(ent)
            k += 1 + nb
        # corner goes to the facet listed first in the user's facet_types (cloud.py:680-688)
        for c, ents in corners.items():
            chosen = sorted(ents, key=lambda e: self.facet_precedence[self.facet_names[e]])[0]
            name = self.facet_names[chosen]
            types[c] = self.facet_types[name]
            facet_nodes_old[name].append(c)
            tag_nodes[chosen].append(c)
        first = np.array([t[0] for t in types])
        self.Ni = int(np.sum(first == "i"))
        self.Nd = int(np.sum(first == "d"))
        self.Nr = int(np.sum(first == "r"))
        self.Nn = int(np.sum(first == "n"))
        return types, coords, facet_nodes_old, tag_nodes

    def _normals(self, types, coords, tag_nodes):
        """Approximate outward normals (cloud.py:698-734): perpendicular to the chord towards the
        nearest node of the same curve, flipped away from the *second* hit of a k=2 query among the
        internal nodes (the reference indexes ``neighbours[0][1]``)."""
        from sklearn.neighbors import BallTree
        normals = {}
        in_coords = coords[np.array([t == "i" for t in types])]
        in_tree = BallTree(in_coords, leaf_size=40, metric="euclidean")
        for tag, ids in tag_nodes.items():
            if self.facet_types[self.facet_names[tag]][0] not in ("n", "r", "p"):
                continue
            assert len(ids) >= 2, " Mesh not fine enough for normal computation "
            f_coords = coords[np.asarray(ids)]
            f_tree = BallTree(f_coords, leaf_size=40, metric="euclidean")
            _, nb_f = f_tree.query(f_coords, k=2)
            _, nb_i = in_tree.query(f_coords, k=2)
            for row, nid in enumerate(ids):
                cur = coords[nid]
                tangent = f_coords[nb_f[row][1]] - cur
                invector = in_coords[nb_i[row][1]] - cur
                normal = np.array([-tangent[1], tangent[0]])
                nrm = np.linalg.norm(normal)
                normals[int(nid)] = -normal / nrm if np.dot(normal, invector) > 0 else normal / nrm
        return normals
