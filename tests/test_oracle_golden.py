"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the hot path
(SURVEY.md §8c).  Each test cites the reference test it transcribes.  No GPU, no libgnnb200."""
import numpy as np
import pytest


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


# ---- GNNGraphs/test/query.jl:49-58 — degree, unweighted ---------------------------------------------
def test_degree_unweighted(oracle):
    s, t = [1, 1, 2, 3], [2, 2, 2, 4]
    assert oracle.degree(s, t, 4, "out").tolist() == [2, 1, 1, 0]
    assert oracle.degree(s, t, 4, "in").tolist() == [0, 3, 0, 1]
    assert oracle.degree(s, t, 4, "both").tolist() == [2, 4, 1, 1]


# ---- GNNGraphs/test/query.jl:71-87 — degree, weighted -----------------------------------------------
def test_degree_weighted(oracle):
    s, t = [1, 1, 2, 3], [2, 2, 2, 4]
    w = np.array([0.1, 2.1, 1.2, 1], dtype=np.float32)
    np.testing.assert_allclose(oracle.degree(s, t, 4, "out", w), [2.2, 1.2, 1.0, 0.0], rtol=1e-6)
    np.testing.assert_allclose(oracle.degree(s, t, 4, "out", 2 * w), [4.4, 2.4, 2.0, 0.0], rtol=1e-6)
    assert oracle.degree(s, t, 4, "out", None).tolist() == [2, 1, 1, 0]


# ---- GNNGraphs/test/transform.jl:1-17 — add_self_loops, A -> A2 (existing loop doubled) ---------------
def test_add_self_loops(oracle):
    A = np.array([[1, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0]])
    A2 = np.array([[2, 1, 0, 0], [0, 1, 1, 0], [0, 0, 1, 1], [1, 0, 0, 1]])
    t, s = np.nonzero(A.T)
    s, t = s + 1, t + 1
    assert (oracle.dense_adjacency(s, t, 4) == A).all()
    s2, t2 = oracle.add_self_loops(s, t, 4)
    assert len(s2) == A2.sum()
    assert (oracle.dense_adjacency(s2, t2, 4) == A2).all()
    # loops are appended AFTER the originals, in node order (transform.jl:17-19)
    assert s2[-4:].tolist() == [1, 2, 3, 4] and t2[-4:].tolist() == [1, 2, 3, 4]
    assert (s2[:-4] == s).all() and (t2[:-4] == t).all()


# ---- GNNlib/test/msgpass.jl:21-26 — isolated node keeps shape, gets 0 for `+` -------------------------
def test_isolated_nodes(oracle):
    x1 = np.random.default_rng(0).random((6, 1)).astype(np.float32)
    s = t = np.arange(1, 6)
    y = oracle.propagate_unfused("+", s, t, 6, x1)
    assert y.shape == (6, 1)
    assert y[5, 0] == 0 and (y[:5] == x1[:5]).all()
    assert oracle.propagate_unfused("mean", s, t, 6, x1)[5, 0] == 0
    assert oracle.propagate_unfused("max", s, t, 6, x1)[5, 0] == -np.inf   # NNlib: typemin (unpinned by reference tests)
    assert oracle.propagate_unfused("min", s, t, 6, x1)[5, 0] == np.inf


def _sprand_graph(n=128, density=0.1, seed=0, weighted=False):
    rng = np.random.default_rng(seed)
    M = (rng.random((n, n)) < density)
    A = M * rng.random((n, n))
    t, s = np.nonzero(A.T)  # column-major findnz order, like GNNGraph(A, graph_type=:coo)
    w = A[s, t]
    return s + 1, t + 1, w, A


# ---- GNNlib/test/msgpass.jl:69-89 — copy_xj + fused and unfused ≈ X * Adj ----------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_copy_xj_sum_matches_dense(oracle, dtype):
    n = 128
    s, t, _, A = _sprand_graph(n)
    Adj = (A > 0).astype(np.float64)
    X = np.random.default_rng(1).random((n, 10)).astype(dtype)       # rows = Julia (10, n)
    ref = (X.astype(np.float64).T @ Adj).T                             # X * Adj
    tol = 1e-6 if dtype == np.float32 else 1e-14
    assert rel(oracle.propagate_unfused("+", s, t, n, X), ref) < tol
    assert rel(oracle.propagate_fused(s, t, n, X), ref) < tol


# ---- GNNlib/test/msgpass.jl:91-116 — e_mul_xj / w_mul_xj ≈ X * A, three ways -------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_weighted_sum_matches_dense(oracle, dtype):
    n = 128
    s, t, w, A = _sprand_graph(n, weighted=True)
    X = np.random.default_rng(2).random((n, 10)).astype(dtype)
    ref = (X.astype(np.float64).T @ A).T
    tol = 1e-6 if dtype == np.float32 else 1e-14
    assert rel(oracle.propagate_unfused("+", s, t, n, X, w), ref) < tol
    assert rel(oracle.propagate_fused(s, t, n, X, w), ref) < tol


def test_duplicate_edges_are_summed(oracle):
    # multi-edges are legal (GNNGraphs/src/gnngraph.jl:36-37): A[i,j] counts them
    s, t = [1, 1, 2], [2, 2, 1]
    X = np.array([[1.0], [10.0]])
    assert oracle.propagate_fused(s, t, 2, X).ravel().tolist() == [10.0, 2.0]
    assert oracle.propagate_unfused("+", s, t, 2, X).ravel().tolist() == [10.0, 2.0]


# ---- GNNlib/test/utils.jl:58-67 — softmax_edge_neighbors ≡ softmax per target segment -----------------
def test_softmax_edge_neighbors(oracle):
    s, t = [1, 2, 3, 4], [5, 5, 6, 6]
    e2 = np.random.default_rng(3).standard_normal((4, 3)).astype(np.float32)   # Julia (3, 4)
    z = oracle.softmax_edge_neighbors(t, 6, e2)
    assert z.shape == e2.shape

    def softmax(a):  # NNlib.softmax(e2[:, 1:2], dims=2): over the edges of the segment
        a = a.astype(np.float64)
        m = np.exp(a - a.max(axis=0, keepdims=True))
        return m / m.sum(axis=0, keepdims=True)

    np.testing.assert_allclose(z[0:2], softmax(e2[0:2]), rtol=1e-6)
    np.testing.assert_allclose(z[2:4], softmax(e2[2:4]), rtol=1e-6)


# ---- GraphNeuralNetworks/test/layers/conv.jl:30-44 — GCN closed form with edge weights ----------------
@pytest.mark.parametrize("fused", [True, False])
def test_gcn_closed_form(oracle, fused):
    s, t = [2, 3, 1, 3, 1, 2], [1, 1, 2, 2, 3, 3]
    w = np.array([1, 2, 3, 4, 5, 6], dtype=np.float32)
    x = np.ones((3, 1), dtype=np.float32)
    d = oracle.degree(s, t, 3, "in", w)
    y, c = oracle.gcn_propagate(s, t, 3, x, w, fused=fused)     # weight = 1, no self loops, no bias
    np.testing.assert_allclose(y[0, 0], w[0] / np.sqrt(d[0] * d[1]) + w[1] / np.sqrt(d[0] * d[2]), rtol=1e-6)
    np.testing.assert_allclose(y[1, 0], w[2] / np.sqrt(d[1] * d[0]) + w[3] / np.sqrt(d[1] * d[2]), rtol=1e-6)
    np.testing.assert_allclose(c, 1 / np.sqrt(d), rtol=1e-7)


# ---- GNNlib/test/msgpass.jl:139-144 (mean) and NNlib semantics for max/min -----------------------------
def test_mean_max_min_against_numpy(oracle):
    rng = np.random.default_rng(4)
    n, E, D = 50, 400, 7
    s = rng.integers(1, n + 1, E)
    t = rng.integers(1, n // 2 + 1, E)        # upper half of the nodes has no in-edges
    x = rng.standard_normal((n, D))
    m = x[s - 1]
    for aggr, fn, empty in (("mean", np.mean, 0.0), ("max", np.max, -np.inf), ("min", np.min, np.inf)):
        got = oracle.propagate_unfused(aggr, s, t, n, x)
        for i in range(n):
            sel = m[t == i + 1]
            exp = fn(sel, axis=0) if len(sel) else np.full(D, empty)
            np.testing.assert_allclose(got[i], exp, rtol=1e-12)


# ---- GAT edge part against an independent numpy restatement ---------------------------
def test_gat_aggregate_against_numpy(oracle):
    rng = np.random.default_rng(5)
    n, E, H, Cc = 20, 90, 3, 4
    s = np.concatenate([rng.integers(1, n + 1, E), np.arange(1, n + 1)])
    t = np.concatenate([rng.integers(1, n + 1, E), np.arange(1, n + 1)])
    Wx = rng.standard_normal((n, H, Cc))
    a = rng.standard_normal((H, 2 * Cc))
    out, alpha = oracle.gat_aggregate(s, t, n, Wx, a, 0.2)
    el = (Wx * a[None, :, :Cc]).sum(-1)
    er = (Wx * a[None, :, Cc:]).sum(-1)
    z = el[t - 1] + er[s - 1]
    u = np.where(z > 0, z, 0.2 * z)
    exp_out = np.zeros_like(Wx)
    for i in range(n):
        sel = np.nonzero(t == i + 1)[0]
        p = np.exp(u[sel] - u[sel].max(0))
        p = p / p.sum(0)
        np.testing.assert_allclose(alpha[sel], p, rtol=1e-10)
        exp_out[i] = (p[:, :, None] * Wx[s[sel] - 1]).sum(0)
    np.testing.assert_allclose(out, exp_out, rtol=1e-10, atol=1e-12)


# ---- CSR index arithmetic: rowptr differences == degree(dir=:in), stable permutation ------------------
def test_csr_index_arithmetic(oracle):
    rng = np.random.default_rng(6)
    n, E = 40, 300
    s = rng.integers(1, n + 1, E)
    t = rng.integers(1, n + 1, E)
    rowptr, col, perm = oracle.csr(t, s, n)
    assert (np.diff(rowptr) == oracle.degree(s, t, n, "in").astype(np.int64)).all()
    assert (t[perm] - 1 == np.repeat(np.arange(n), np.diff(rowptr))).all()
    assert (col == s[perm] - 1).all()
    for i in range(n):
        seg = perm[rowptr[i]:rowptr[i + 1]]
        assert (np.diff(seg) > 0).all()       # stable: COO order inside a row


def test_rmat_deterministic_and_in_range(oracle):
    s, t = oracle.rmat(1000, 5000, 17)
    s2, t2 = oracle.rmat(1000, 5000, 17)
    assert (s == s2).all() and (t == t2).all()
    assert s.min() >= 1 and s.max() <= 1000 and t.min() >= 1 and t.max() <= 1000
    # skew: the low-id quadrant is the heavy one (a = 0.57)
    assert (s <= 500).mean() > 0.6 and (t <= 500).mean() > 0.6
    # known-answer prefix (pins the generator across CPU and GPU; regenerate only with the generator)
    assert s[:4].tolist() == RMAT_S4 and t[:4].tolist() == RMAT_T4


RMAT_S4 = [35, 34, 10, 1]
RMAT_T4 = [829, 337, 130, 27]


def test_csc_build_and_products(oracle):
    # sparse(s,t,w) sums duplicates; xj*A and its pullback Δ*A' against dense linear algebra
    rng = np.random.default_rng(8)
    n, E, D = 30, 200, 5
    s = rng.integers(1, n + 1, E); t = rng.integers(1, n + 1, E)
    w = rng.random(E)
    A = oracle.dense_adjacency(s, t, n, w)
    csc = oracle.csc_build(s, t, n, w, np.float64)
    assert len(csc[1]) == (A != 0).sum()
    x = rng.standard_normal((n, D))
    np.testing.assert_allclose(oracle.dense_times_csc(x, csc), (x.T @ A).T, rtol=1e-12)
    np.testing.assert_allclose(oracle.dense_times_csc(x, csc, transposed=True), (x.T @ A.T).T, rtol=1e-12)
    cu = oracle.csc_build(s, t, n, None, np.float32)      # unweighted: integer counts
    assert (cu[2] == np.round(cu[2])).all() and cu[2].sum() == E


# ---- tests/golden/reference_known_answers.json: every transcribed known answer, replayed through the oracle -------
def test_known_answer_file(oracle):
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_known_answers.json")) as f:
        G = json.load(f)
    d = G["degree"]
    for key in ("out", "in", "both"):
        assert oracle.degree(d["s"], d["t"], d["n"], key).tolist() == d[key], d["source"]
    np.testing.assert_allclose(oracle.degree(d["s"], d["t"], d["n"], "out", np.asarray(d["w"], np.float32)),
                               d["out_weighted"], rtol=1e-6)
    a = G["add_self_loops"]
    A, A2 = np.asarray(a["A"]), np.asarray(a["A2"])
    t, s = np.nonzero(A.T)
    s2, t2 = oracle.add_self_loops(s + 1, t + 1, 4)
    assert (oracle.dense_adjacency(s2, t2, 4) == A2).all(), a["source"]
    sm = G["softmax_edge_neighbors"]
    e = np.random.default_rng(0).standard_normal((4, 3))
    got = oracle.softmax_edge_neighbors(sm["t"], sm["n"], e)
    for lo in (0, 2):
        ex = np.exp(e[lo:lo + 2] - e[lo:lo + 2].max(0))
        assert rel(got[lo:lo + 2], ex / ex.sum(0)) < 1e-14, sm["source"]
    c = G["gcn_closed_form"]
    w = np.asarray(c["w"], dtype=np.float64)
    y, _ = oracle.gcn_propagate(c["s"], c["t"], c["n"], np.ones((3, 1)), w)
    dg = oracle.degree(c["s"], c["t"], c["n"], "in", w, dtype=np.float64)
    assert abs(y[0, 0] - (w[0] / np.sqrt(dg[0] * dg[1]) + w[1] / np.sqrt(dg[0] * dg[2]))) < 1e-12, c["source"]
    assert abs(y[1, 0] - (w[2] / np.sqrt(dg[1] * dg[0]) + w[3] / np.sqrt(dg[1] * dg[2]))) < 1e-12
    b = G["to_bidirected"]                                   # remove_multi_edges(mean) of [s;t],[t;s] in (s,t) order
    s, t = np.asarray(b["s"] + b["t"]), np.asarray(b["t"] + b["s"])
    enc = (s - 1) * 4 + t
    perm = np.argsort(enc, kind="stable")
    head = np.concatenate([[True], enc[perm][1:] > enc[perm][:-1]])
    seg = np.cumsum(head)
    assert s[perm][head].tolist() == b["s2"] and t[perm][head].tolist() == b["t2"], b["source"]
    for key, out in (("w", "w2"), ("e", "e2")):
        v = np.asarray(b[key] + b[key])[perm]
        assert oracle.scatter("mean", v[:, None], seg, int(seg[-1]))[:, 0].tolist() == b[out]
    for name in ("symmetric_graph", "asymmetric_graph"):
        gr = G[name]
        A = oracle.dense_adjacency(gr["s"], gr["t"], 4)
        assert A.astype(int).tolist() == gr.get("adj", gr.get("adj_out")), gr["source"]
