"""The warp fast path (magnetizer_b200/csrc/mag_rk_warp.cuh) never stores ghost values: the
mirror conditions of impose_bc (gutsdynamo.f90:55-86: Br, Bp odd about the two Dirichlet
points, alp_m even) are folded into the stencil weights of the three points next to each end
(fold_boundary_weights), Br/Bp ghost entries read 0 and three alp_m ghost entries per end
hold the mirror values.  This is a numpy restatement of that algebra, checked against the
straightforward evaluation with explicit ghost zones, plus the lane layout arithmetic.
(The CUDA code itself is checked against the oracle by the -m gpu tests.)"""
import numpy as np


def fold(w, w0a, d, left, rk=1.0):
    """fold_boundary_weights for the point at distance d from the Dirichlet point.
    w[k], k = 0..6 <-> offsets -3..+3; returns (w, w0a) folded."""
    w = w.copy()
    wc = w[3]
    for o in range(-3, -d):                      # ghost offsets, inward positive
        io = 3 + o if left else 3 - o
        om = -2 * d - o
        im = 3 + om if left else 3 - om
        if d == 0:
            w[io] += w[im]; w[im] = 0.0
        elif om == 0:
            wc -= w[io]; w0a += rk * w[io]; w[io] = 0.0
        else:
            w[im] -= w[io]; w[io] *= 2.0
    w[3] = wc
    return w, w0a


def test_folded_weights_reproduce_the_mirror_conditions():
    rng = np.random.default_rng(7)
    n = 40                                       # live points p = 0..n-1 (Dirichlet points at both ends)
    for trial in range(20):
        W = rng.normal(size=(n, 7))              # per-point stencil weights (Br, Bp and, times R_kappa = 1, alp_m)
        W0A = rng.normal(size=n)                 # centre weight of the alp_m equation
        B = rng.normal(size=n); B[0] = B[-1] = 0.0       # Dirichlet
        A = rng.normal(size=n)

        def ghosted(f, sigma):                   # explicit ghost zones of width 3
            g = np.empty(n + 6)
            g[3:n + 3] = f
            for k in (1, 2, 3):
                g[3 - k] = sigma * f[k]
                g[n + 2 + k] = sigma * f[n - 1 - k]
            return g

        gB, gA = ghosted(B, -1.0), ghosted(A, +1.0)
        # reference: L(B)_p and the alp_m chain with real ghosts
        LB = np.array([W[p] @ gB[p:p + 7] for p in range(n)])
        LA = np.array([np.delete(W[p], 3) @ np.delete(gA[p:p + 7], 3) + W0A[p] * A[p] for p in range(n)])
        # folded: weights modified at d = 0, 1, 2 from each end; B ghost entries 0, alp_m ghost entries mirrored
        Wf, W0Af = W.copy(), W0A.copy()
        for d in range(3):
            Wf[d], W0Af[d] = fold(W[d], W0A[d], d, True)
            Wf[n - 1 - d], W0Af[n - 1 - d] = fold(W[n - 1 - d], W0A[n - 1 - d], d, False)
        zB = np.zeros(n + 6); zB[3:n + 3] = B    # inert slots read 0
        LBf = np.array([Wf[p] @ zB[p:p + 7] for p in range(n)])
        LAf = np.array([np.delete(Wf[p], 3) @ np.delete(gA[p:p + 7], 3) + W0Af[p] * A[p] for p in range(n)])
        # Br, Bp: exact away from the Dirichlet points; AT them the folded form gives w_c * 0 = 0,
        # i.e. the state stays 0 without being reset (its pre-reset value is never used)
        assert np.allclose(LBf[1:-1], LB[1:-1], rtol=1e-12, atol=1e-12)
        assert LBf[0] == 0.0 and LBf[-1] == 0.0
        assert np.allclose(LAf, LA, rtol=1e-12, atol=1e-12)
        # only three alp_m ghost entries per end carry a non-zero weight
        for p in range(n):
            for k in range(7):
                q = p + k - 3
                if q < -3 or q > n + 2:
                    assert Wf[p, k] == 0.0


def warp_layout(nx):
    """warp_ppl / tm / jm of mag_rk_warp.cuh (WARP_J0 = 3)."""
    P = nx - 7
    for ppl in (4, 5, 6, 8, 11):
        if P + 3 + 3 < 32 * ppl:
            return ppl, (P + 3) // ppl, (P + 3) % ppl
    return 0, 0, 0


def test_warp_layout_covers_every_grid_up_to_352_points():
    for nx in range(64, 353):
        ppl, tm, jm = warp_layout(nx)
        P = nx - 7
        assert ppl in (4, 5, 6, 8, 11)
        assert ppl * tm + jm == P + 3 and tm >= 2            # the outer Dirichlet point has a lane, lanes 0, 1 and tm-1, tm differ
        assert P + 3 + 3 <= 32 * ppl - 1                     # its three ghost slots still fall into the warp
        # the fix-ups of the outer end stay inside the windows of lanes tm and tm-1
        for k in (1, 2, 3):
            assert 0 <= 3 + jm - k and (3 + jm + k <= ppl + 5 or jm + k >= ppl)
    assert warp_layout(107) == (4, 25, 3)                    # default grid: the last live point fills its lane
    assert warp_layout(353)[0] == 0
