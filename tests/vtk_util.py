"""Helpers for the tests of the ported driver's output files (host/src/VTKDatExport.cpp vs src/VTKDatExport.cpp)."""
import re

import numpy as np


def read_solution_vtk(path):
    """Rows (x, y, z, u, v, w, p) of the binary legacy VTK file the writers produce, in file order."""
    blob = open(path, "rb").read()
    n = int(re.search(rb"POINTS (\d+) double\n", blob).group(1))
    off = re.search(rb"POINTS \d+ double\n", blob).end()
    cols = [np.frombuffer(blob[off:off + 24 * n], dtype=">f8").reshape(n, 3)]
    off += 24 * n
    for name in "uvwp":
        m = re.search(rb"SCALARS " + name.encode() + rb" double 1\nLOOKUP_TABLE default\n", blob[off:])
        off += m.end()
        cols.append(np.frombuffer(blob[off:off + 8 * n], dtype=">f8").reshape(n, 1))
        off += 8 * n
    assert off == len(blob)
    return np.hstack(cols).astype(np.float64)


def sorted_rows(rows):
    key = np.round(rows[:, :3] * 1e9).astype(np.int64)
    order = np.lexsort((rows[:, 6], rows[:, 5], rows[:, 4], rows[:, 3], key[:, 2], key[:, 1], key[:, 0]))
    return rows[order]


def block_first(n, parts):
    return [n // parts * r + min(r, n % parts) for r in range(parts + 1)]


def check_multi_rank_solution(path, golden_path, N, lo, h, Py, Pz):
    """A file written by Py x Pz ranks against the single-rank golden: the same multiset of rows (the writers list the
    points rank by rank -- the MPI-IO offsets of src/VTKDatExport.cpp:219-311 -- so only the order differs), and the
    points really come in rank order: rank = y_rank * Pz + z_rank owns the y rows / z planes of its blocks
    (src/Constants.cpp:78-94)."""
    got, ref = read_solution_vtk(path), read_solution_vtk(golden_path)
    assert got.shape == ref.shape
    a, b = sorted_rows(got), sorted_rows(ref)
    scale = np.maximum(np.max(np.abs(b), axis=0), 1e-6)
    assert np.max(np.abs(a - b) / scale) <= 1e-10
    ys, zs = block_first(N[1], Py), block_first(N[2], Pz)
    j = np.round((got[:, 1] - lo[1]) / h[1]).astype(int)
    k = np.round((got[:, 2] - lo[2]) / h[2]).astype(int)
    rank = (np.searchsorted(ys, j, side="right") - 1) * Pz + (np.searchsorted(zs, k, side="right") - 1)
    assert np.all(np.diff(rank) >= 0), "points are not listed rank by rank"
    assert len(np.unique(rank)) == Py * Pz
