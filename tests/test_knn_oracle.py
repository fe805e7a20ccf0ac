"""CPU tests of oracle/knn_oracle.py (builder-defined semantics of BASELINE config 4): the two candidate routes agree,
the float32-ordered index lists are pinned against scipy.spatial.cKDTree.query, the blend's gradient is checked by
central differences, and the cell function reproduces a hand-computed case."""
import numpy as np
import torch

from oracle import knn_oracle as KO


def _scene(P=4000, N=600, seed=0, radius=0.2):
    rng = np.random.default_rng(seed)
    lo, hi = np.array([-1.0, -0.5, 0.0], np.float32), np.array([1.0, 1.0, 1.5], np.float32)
    xyz = (lo + (hi - lo) * rng.random((P, 3), dtype=np.float32)).astype(np.float32)
    p = (lo - 0.3 + (hi - lo + 0.6) * rng.random((N, 3), dtype=np.float32)).astype(np.float32)   # some queries outside
    return xyz, p, lo, hi, radius


def test_ball_route_equals_bruteforce_and_handles_sparse_neighbourhoods():
    xyz, p, lo, hi, r = _scene()
    i1, d1 = KO.knn_query(p, xyz, r)
    i2, d2 = KO.knn_query_bruteforce(p, xyz, r)
    assert np.array_equal(i1, i2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
    n_found = (i1 >= 0).sum(1)
    assert n_found.min() == 0 and n_found.max() == 8 and ((n_found > 0) & (n_found < 8)).any()   # empty, partial and full lists
    full = i1[n_found == 8]
    assert (np.diff(d1[n_found == 8], axis=1) >= 0).all() and len(np.unique(full[0])) == 8     # ascending, distinct


def test_ties_break_by_point_index():
    xyz = np.zeros((12, 3), np.float32)
    xyz[:, 0] = [0.1, -0.1, 0.1, -0.1, 0.1, -0.1, 0.1, -0.1, 0.1, -0.1, 0.05, 0.3]   # ten points at the same distance
    idx, d2 = KO.knn_query_bruteforce(np.zeros((1, 3), np.float32), xyz, 0.2)
    assert idx[0].tolist() == [10, 0, 1, 2, 3, 4, 5, 6]


def test_indices_pinned_against_ckdtree():
    xyz, p, lo, hi, r = _scene(P=20000, N=3000, seed=1, radius=0.12)
    idx, _ = KO.knn_query(p, xyz, r)
    n_bad, worst = KO.pin_against_ckdtree(p, xyz, r, idx)
    # float64 vs float32 ordering may differ only between neighbours whose distances agree to float32 rounding
    assert n_bad <= 3 and worst < 2e-6, (n_bad, worst)


def test_cell_function_and_grouping():
    lo, inv_h, dims = [0.0, 0.0, 0.0], np.float32(1.0 / 0.25), (4, 3, 2)
    xyz = np.array([[0.0, 0.0, 0.0], [0.26, 0.1, 0.3], [0.99, 0.74, 0.49], [5.0, -1.0, 0.25]], np.float32)
    c = KO.cell_coords(xyz, lo, inv_h, dims)
    assert c.tolist() == [[0, 0, 0], [1, 0, 1], [3, 2, 1], [3, 0, 1]]
    start, order = KO.build_cells(xyz, lo, inv_h, dims)
    assert start[-1] == 4 and sorted(order.tolist()) == [0, 1, 2, 3]
    cid = (c[:, 2] * 3 + c[:, 1]) * 4 + c[:, 0]
    for cell in np.unique(cid):
        assert set(order[start[cell]:start[cell + 1]].tolist()) == set(np.nonzero(cid == cell)[0].tolist())


def test_blend_weights_and_gradient():
    xyz, p, lo, hi, r = _scene(P=150, N=12, seed=2, radius=0.6)
    idx, d2 = KO.knn_query(p, xyz, r)
    feat = torch.randn(150, 4, dtype=torch.float64)
    X, I = torch.from_numpy(xyz).double(), torch.from_numpy(idx)

    def f(pp, ff):
        # float64 copy of aggregate() for the finite-difference check (p.float() inside aggregate would lose the step)
        valid = I >= 0
        j = I.clamp(min=0).long()
        d = pp[:, None, :] - X[j]
        dd = (d * d).sum(-1)
        w = torch.where(valid, 1.0 / (dd + 1e-6), torch.zeros_like(dd))
        W = w.sum(1, keepdim=True)
        wn = torch.where(W > 0, w / torch.where(W > 0, W, torch.ones_like(W)), torch.zeros_like(w))
        return (wn[..., None] * ff[j]).sum(1)

    pt = torch.from_numpy(p).double().requires_grad_(True)
    ft = feat.clone().requires_grad_(True)
    assert torch.autograd.gradcheck(f, (pt, ft), eps=1e-7, atol=1e-5, rtol=1e-4, nondet_tol=0.0)
    # the float32 restatement agrees with it and gives a convex combination
    out32 = KO.aggregate(torch.from_numpy(p), torch.from_numpy(xyz), feat.float(), I, 1e-6)
    assert torch.allclose(out32.double(), f(pt, ft).detach(), atol=1e-5, rtol=1e-4)
    one = KO.aggregate(torch.from_numpy(p), torch.from_numpy(xyz), torch.ones(150, 4), I, 1e-6)
    has = torch.from_numpy((idx >= 0).any(1))
    assert torch.allclose(one[has], torch.ones_like(one[has]), atol=1e-6) and bool((one[~has] == 0).all())
