"""Shared test helpers: seeded data and the parity comparator of SURVEY.md §8c."""
import numpy as np

REL_TOL = 1e-4  # north_star: distances within 1e-4 relative; ids identical up to ties inside that band


def gen(n, d, seed, dist="uniform"):
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        return rng.random((n, d), dtype=np.float32)
    centers = rng.random((64, d), dtype=np.float32)
    lab = rng.integers(0, 64, n)
    return (centers[lab] + 0.1 * rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


def close(a, b, tol=REL_TOL):
    return abs(a - b) <= tol * max(abs(a), abs(b), 1e-30) + 1e-7


def assert_same_results(got_ids, got_d, got_n, want_ids, want_d, want_n, what="", tol=REL_TOL, allow_boundary=True):
    """Position-wise id equality, tolerating permutations inside groups whose oracle distances are within
    `tol` relative (fp32 summation order differs: the reference's value is GCC's SSE2 lane order).  At the
    cut-off position an element may be swapped for an equally-distant one.  Returns the exact-match rate."""
    got_ids, want_ids = np.atleast_2d(got_ids), np.atleast_2d(want_ids)
    got_d, want_d = np.atleast_2d(got_d), np.atleast_2d(want_d)
    exact = 0
    for q in range(want_ids.shape[0]):
        n = int(want_n[q])
        assert int(got_n[q]) == n, "%s query %d: count %d != %d" % (what, q, got_n[q], n)
        g, w = got_ids[q, :n], want_ids[q, :n]
        gd, wd = got_d[q, :n], want_d[q, :n]
        for i in range(n):
            assert close(gd[i], wd[i], tol), "%s query %d pos %d: dist %r vs %r" % (what, q, i, gd[i], wd[i])
        if np.array_equal(g, w):
            exact += 1
            continue
        for i in range(n):
            if g[i] == w[i]:
                continue
            # g[i] must appear in the oracle list at a position whose distance ties with position i ...
            j = np.nonzero(w == g[i])[0]
            if len(j):
                assert close(wd[j[0]], wd[i], tol), "%s query %d pos %d: id %d out of order beyond ties" % (what, q, i, g[i])
            else:
                # ... or be a boundary swap: same distance as the oracle's last kept entry
                assert allow_boundary and close(gd[i], wd[n - 1], tol), \
                    "%s query %d pos %d: id %d not in oracle result" % (what, q, i, g[i])
    return exact / max(1, want_ids.shape[0])


def recall(got_ids, truth_ids, k):
    hit = 0
    for g, t in zip(got_ids, truth_ids):
        hit += len(set(g[:k].tolist()) & set(t[:k].tolist()))
    return hit / (k * len(truth_ids))


def exact_topk(X, Q, k, metric="l2"):
    """float64 ground truth (ids only; used for recall, never for parity)."""
    X64, Q64 = X.astype(np.float64), Q.astype(np.float64)
    if metric == "l2":
        d = (Q64 ** 2).sum(1)[:, None] - 2 * Q64 @ X64.T + (X64 ** 2).sum(1)[None, :]
    elif metric == "ip":
        d = -(Q64 @ X64.T)
    else:
        d = 1 - Q64 @ X64.T
    idx = np.argsort(d, axis=1, kind="stable")[:, :k]
    return idx
