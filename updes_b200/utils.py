"""Small host utilities of the reference surface that its demo scripts call next to the solver (``updes/utils.py:152-246``,
``updes/operators.py:777-779``).  None of them is on the hot path; they exist so a reference script keeps running after
``import updes_b200 as updes``."""
from __future__ import annotations

import os
import random

import numpy as np


def random_name(length=5):
    """A string of random digits to label a run (utils.py:152-157)."""
    return "".join(str(random.randint(0, 9)) for _ in range(length))


def make_dir(path):
    """Create a directory if it does not exist (utils.py:160-163)."""
    if not os.path.exists(path):
        os.mkdir(path)


def print_line_by_line(dictionary):
    """utils.py:26-28"""
    for k, v in dictionary.items():
        print("\t", k, ":", v)


def dot_vec(a, b):
    """Row-wise dot products of two (R, d) arrays (operators.py:778)."""
    return np.einsum("ij,ij->i", np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))


def dot_mat(J, v):
    """Row-wise matrix-vector products: (R, d, d) with (R, d) (operators.py:779)."""
    return np.einsum("ijk,ik->ij", np.asarray(J, dtype=np.float64), np.asarray(v, dtype=np.float64))


def RK4(fun, t_span, y0, *args, t_eval=None, subdivisions=1, **kwargs):
    """Fixed-step classical Runge-Kutta with ``subdivisions`` sub-steps per evaluation interval (utils.py:198-246).
    Returns the solution at the times of ``t_eval``, shape (len(t_eval), ...)."""
    if t_eval is None:
        if t_span[0] is None:
            raise Warning("t_span[0] is None. Setting t_span[0] to 0.")
        if t_span[1] is None:
            raise ValueError("t_span[1] must be provided if t_eval is not.")
        t_eval = np.array(t_span, dtype=np.float64)
    t_eval = np.asarray(t_eval, dtype=np.float64)
    hs = t_eval[1:] - t_eval[:-1]
    t_ = t_eval[:-1, None] + np.arange(subdivisions)[None, :] * hs[:, None] / subdivisions
    t_solve = np.concatenate([t_.flatten(), t_eval[-1:]])
    t_prev, y = t_solve[0], np.asarray(y0, dtype=np.float64)
    ys = []
    for t in t_solve:                           # the first step has h = 0 and returns y0, as the reference's scan does
        h = t - t_prev
        k1 = h * fun(t_prev, y, *args)
        k2 = h * fun(t_prev + h / 2.0, y + k1 / 2.0, *args)
        k3 = h * fun(t_prev + h / 2.0, y + k2 / 2.0, *args)
        k4 = h * fun(t + h, y + k3, *args)
        y = y + (k1 + 2 * k2 + 2 * k3 + k4) / 6.0
        t_prev = t
        ys.append(y)
    return np.stack(ys)[np.arange(0, t_solve.size, subdivisions)]


def plot(*args, ax=None, figsize=(6, 3.5), x_label=None, y_label=None, title=None, x_scale="linear", y_scale="linear",
         xlim=None, ylim=None, **kwargs):
    """Line plot on a (new) matplotlib axis (utils.py:164-185).  Needs matplotlib, which the solver itself does not."""
    try:
        import matplotlib.pyplot as plt
    except ImportError as e:
        raise ImportError("updes_b200.plot needs matplotlib, which is not installed; the solver itself does not") from e
    if ax is None:
        _, ax = plt.subplots(1, 1, figsize=figsize)
    if x_label:
        ax.set_xlabel(x_label)
    if y_label:
        ax.set_ylabel(y_label)
    if title:
        ax.set_title(title)
    ax.plot(*args, **kwargs)
    ax.set_xscale(x_scale)
    ax.set_yscale(y_scale)
    if "label" in kwargs:
        ax.legend()
    if ylim:
        ax.set_ylim(ylim)
    if xlim:
        ax.set_xlim(xlim)
    plt.tight_layout()
    return ax


def dataloader(array, batch_size, key):
    """Shuffled mini-batches of the rows of ``array`` (utils.py:188-197): a permutation drawn from ``key`` (an integer seed
    or a numpy Generator here; a jax PRNG key in the reference, whose random stream cannot be reproduced without JAX), then
    consecutive slices of it while a whole batch BEFORE the end of the data remains -- the reference's loop condition
    ``end < dataset_size``, which never yields the last batch."""
    array = np.asarray(array)
    n = array.shape[0]
    rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(key)
    perm = rng.permutation(n)
    start, end = 0, batch_size
    while end < n:
        yield array[perm[start:end]]
        start, end = end, end + batch_size
