"""Helpers shared by the parity tests: identity-independent comparison of particle state.

The product keeps particles in cell order, the reference in its own (hole-filling) order, so states are compared
through particle identity: locals by their lattice/upload index (product: `tag`, oracle: `uid`, which the tests set
to the same numbering before the first step), ghosts by (identity, exact coordinates)."""
import numpy as np


def rows_sorted(*cols):
    """Stack columns into rows and sort them lexicographically (first column most significant)."""
    a = np.stack([np.asarray(c) for c in cols], axis=1)
    order = np.lexsort(tuple(a[:, k] for k in range(a.shape[1] - 1, -1, -1)))
    return a[order]


def f2i(x):
    """float64 -> int64 bit pattern with -0.0 folded onto +0.0 (they compare equal numerically)."""
    x = np.asarray(x, np.float64) + 0.0
    return x.view(np.int64)


def by_id(ids, values):
    """Reorder `values` (first axis) so that row k belongs to identity k."""
    order = np.argsort(ids, kind="stable")
    assert np.array_equal(np.asarray(ids)[order], np.arange(len(ids))), "identities are not a permutation"
    return np.asarray(values)[order]


def neighbor_rows(ids, pos, numneighs, neigh, nlocal):
    """All (id_i, id_j, xj, yj, zj) rows of a set of neighbour lists, sorted -> order-independent representation."""
    ids = np.asarray(ids)
    nn = np.asarray(numneighs)[:nlocal]
    total = int(nn.sum())
    ii = np.repeat(np.arange(nlocal), nn)
    kk = np.arange(total) - np.repeat(np.cumsum(nn) - nn, nn)
    jj = np.asarray(neigh)[ii, kk]
    pj = f2i(pos[jj])
    return rows_sorted(ids[ii].astype(np.int64), ids[jj].astype(np.int64), pj[:, 0], pj[:, 1], pj[:, 2])


def rel_err_force(f, f_ref):
    """max-norm relative error ||f - f_ref||_inf / ||f_ref||_inf (per-particle |f_i| can vanish by symmetry)."""
    scale = np.abs(f_ref).max()
    return float(np.abs(f - f_ref).max() / scale) if scale > 0 else float(np.abs(f - f_ref).max())
