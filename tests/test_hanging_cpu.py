"""Hanging-node bookkeeping (CPU): the compressed description handed to the C ABI (parent face /
edge nodes in the cell list + refinement mask + 1-D interpolation per direction) reproduces the
plain constraint expansion used by the oracle."""
import numpy as np
import pytest

from adaflo_b200.mesh import (HN_EDGE_X, HN_FACE_X, HN_SUBCELL_X, RefinedBoxMesh, gauss_lobatto_01, lagrange_values)


def interpolation_matrix(n1, sub, iso_q1=False):
    """values of the parent's 1-D basis at the child's nodes (child 0: [0,1/2], child 1: [1/2,1])"""
    if iso_q1:   # hats on n1 - 1 sub-intervals, equidistant nodes
        x = 0.5 * np.arange(n1) / (n1 - 1) + 0.5 * sub
        return np.maximum(0.0, 1.0 - np.abs(x[:, None] * (n1 - 1) - np.arange(n1)[None, :]))
    gl = gauss_lobatto_01(n1)
    return lagrange_values(gl, 0.5 * gl + 0.5 * sub)[0]


def resolve_compressed(mask, vals, n1, iso_q1=False):
    """vals: [n1,n1,n1] indexed [z][y][x] holding the slot values; returns the cell-local nodal values.
    This is the algorithm the CUDA kernels implement (direction by direction, each line at most once)."""
    v = vals.copy()
    if mask == 0:
        return v
    sub = [(mask >> d) & 1 for d in range(3)]
    face = [bool(mask & (HN_FACE_X << d)) for d in range(3)]
    edge = [bool(mask & (HN_EDGE_X << d)) for d in range(3)]
    outer = [0 if sub[d] == 0 else n1 - 1 for d in range(3)]
    for d in range(3):                      # interpolate along direction d
        e, f = (d + 1) % 3, (d + 2) % 3
        M = interpolation_matrix(n1, sub[d], iso_q1)
        for ie in range(n1):
            for jf in range(n1):
                need = (face[e] and ie == outer[e]) or (face[f] and jf == outer[f]) or \
                       (edge[d] and ie == outer[e] and jf == outer[f])
                if not need:
                    continue
                sl = [None, None, None]
                sl[d], sl[e], sl[f] = slice(None), ie, jf
                key = (sl[2], sl[1], sl[0])           # array is [z][y][x]
                v[key] = M @ v[key]
    return v


@pytest.mark.parametrize("refine", [((1, 3), (0, 2), (1, 2)), ((0, 1), (0, 1), (0, 1)), ((1, 2), (1, 2), (0, 3)), ((0, 2), (1, 3), (1, 3))])
def test_compressed_scheme_equals_plain_expansion(refine):
    m = RefinedBoxMesh((3, 3, 3), refine)
    rng = np.random.default_rng(3)
    for sp, lists, n1 in ((m.u, m.u_lists, 3), (m.p, m.p_lists, 2)):
        ptr, idx, w = lists
        vec = rng.uniform(-1, 1, sp.n_owned * sp.n_comp)
        plain = np.add.reduceat(w * vec[idx], ptr[:-1]).reshape(m.n_cells, sp.n_comp, n1, n1, n1)
        for c in range(m.n_cells):
            for comp in range(sp.n_comp):
                slots = vec[sp.cell_nodes[c].astype(np.int64) * sp.n_comp + comp].reshape(n1, n1, n1)
                got = resolve_compressed(int(sp.hanging_mask[c]), slots, n1)
                assert np.abs(got - plain[c, comp]).max() < 1e-14, (c, comp, hex(int(sp.hanging_mask[c])))
        assert (sp.hanging_mask != 0).sum() > 0


def test_expansion_reproduces_the_fe_space():
    """a function of the coarse Q2 space is reproduced at every cell-local node, hanging or not"""
    m = RefinedBoxMesh((3, 2, 2), ((1, 2), (0, 2), (0, 1)))
    f = lambda x: 1 + x[:, 0] * x[:, 1] + 2 * x[:, 2] ** 2 - x[:, 0] * x[:, 2] ** 2
    ptr, idx, w = m.u_lists
    vec = np.repeat(f(m.u.coords)[:, None], 3, axis=1).reshape(-1)
    loc = np.add.reduceat(w * vec[idx], ptr[:-1]).reshape(m.n_cells, 3, 27)
    a = np.arange(3)
    lz, ly, lx = np.meshgrid(a, a, a, indexing="ij")
    xi = np.stack([lx.reshape(-1), ly.reshape(-1), lz.reshape(-1)], axis=1) * 0.5
    for c in range(m.n_cells):
        x = m.origin[c][None, :] + m.h_cell[c][None, :] * xi
        assert np.abs(loc[c, 0] - f(x)).max() < 1e-13


def test_dof_counts():
    m = RefinedBoxMesh((2, 2, 2), ((0, 1), (0, 1), (0, 1)))
    assert m.n_cells == 7 + 8
    # Q2: 5^3 coarse nodes + fine nodes of the refined cell that are not hanging: the refined cell's 5^3 lattice
    # minus its 3^3 coarse nodes minus the hanging ones on the 3 interior-facing faces
    assert m.u.n_owned == 125 + (125 - 27) - 3 * (25 - 9) + 3 * (5 - 3) - 0


def test_level_set_space_with_hanging_nodes():
    m = RefinedBoxMesh((3, 3, 2), ((1, 3), (0, 2), (0, 1)))
    sp = m.build_ls_space(3)
    ptr, idx, w = m.ls_lists
    rng = np.random.default_rng(5)
    vec = rng.uniform(-1, 1, sp.n_owned)
    plain = np.add.reduceat(w * vec[idx], ptr[:-1]).reshape(m.n_cells, 4, 4, 4)
    for c in range(m.n_cells):
        slots = vec[sp.cell_nodes[c].astype(np.int64)].reshape(4, 4, 4)
        got = resolve_compressed(int(sp.hanging_mask[c]), slots, 4, iso_q1=True)
        assert np.abs(got - plain[c]).max() < 1e-14
    assert (sp.hanging_mask != 0).sum() > 0
