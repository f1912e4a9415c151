"""Krylov oracle (ORACLE -- test infrastructure only).

Restates what HYMLS::BaseSolver hands to Belos (src/HYMLS_BaseSolver.cpp:31-139,309-359).
Belos itself (Trilinos, un-vendored, version unpinned: cmake/common.cmake:79) is not in the
reference tree; the algorithm restated here is Belos' published BlockGmresSolMgr / BlockCGSolMgr
behaviour for block size 1:
  * GMRES(m): Arnoldi with two passes of classical Gram-Schmidt (Belos ICGS/DGKS), Givens
    least squares, restart length "Num Blocks" (default 300), "Maximum Restarts" (20),
    "Maximum Iterations" (1000), "Convergence Tolerance" (1e-8);
    implicit residual test on the Givens residual scaled by
    "Implicit Residual Scaling" (default: norm of the (left-)preconditioned initial residual),
    optional "Explicit Residual Test" scaled by "Explicit Residual Scaling"
    (default: norm of initial residual);
  * right preconditioning: A M^-1 y = b - A x0, x = x0 + M^-1 y; left: M^-1 A x = M^-1 b;
  * CG: standard preconditioned CG, test ||r||_2 / ||r0||_2 <= tol.
Iteration counts are pinned only through the reference's integration-test upper bounds.
"""
import numpy as np


def _scale(kind, b, r0, pr0):
    if kind == "Norm of RHS":
        return np.linalg.norm(b)
    if kind == "Norm of Initial Residual":
        return np.linalg.norm(r0)
    if kind == "Norm of Preconditioned Initial Residual":
        return np.linalg.norm(pr0)
    if kind == "None":
        return 1.0
    raise ValueError(kind)


def gmres(apply_A, b, x0, apply_M=None, side="Right", tol=1e-8, max_iters=1000,
          num_blocks=300, max_restarts=20, explicit_test=False,
          imp_scaling="Norm of Preconditioned Initial Residual",
          exp_scaling="Norm of Initial Residual", dot=None):
    """returns x, iters, converged, history (relative implicit residuals per iteration)"""
    b = np.asarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, copy=True)
    n = len(b)
    left = apply_M is not None and side == "Left"
    right = apply_M is not None and side == "Right"
    dotf = (lambda V, w: V @ w) if dot is None else dot

    def prec_res(r):
        return apply_M(r) if left else r

    r0_true = b - apply_A(x)
    pr0 = prec_res(r0_true)
    imp_scale = _scale(imp_scaling, b, r0_true, pr0)
    exp_scale = _scale(exp_scaling, b, r0_true, pr0)
    if imp_scale == 0.0:
        imp_scale = 1.0
    if exp_scale == 0.0:
        exp_scale = 1.0
    history = []
    iters = 0
    converged = False
    r = pr0
    for restart in range(max_restarts + 1):
        beta = np.linalg.norm(r)
        history.append(beta / imp_scale) if restart == 0 else None
        if beta / imp_scale <= tol and not explicit_test:
            converged = True
            break
        m = num_blocks
        V = np.zeros((m + 1, n))
        H = np.zeros((m + 1, m))
        cs = np.zeros(m)
        sn = np.zeros(m)
        g = np.zeros(m + 1)
        g[0] = beta
        V[0] = r / beta
        k_done = 0
        stop = False
        for k in range(m):
            if iters >= max_iters:
                stop = True
                break
            vk = V[k]
            if right:
                w = apply_A(apply_M(vk))
            elif left:
                w = apply_M(apply_A(vk))
            else:
                w = apply_A(vk)
            # two passes of classical Gram-Schmidt
            h = dotf(V[:k + 1], w)
            w = w - V[:k + 1].T @ h
            h2 = dotf(V[:k + 1], w)
            w = w - V[:k + 1].T @ h2
            h = h + h2
            hn = np.linalg.norm(w)
            H[:k + 1, k] = h
            H[k + 1, k] = hn
            if hn > 0:
                V[k + 1] = w / hn
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k])
            cs[k] = H[k, k] / d
            sn[k] = H[k + 1, k] / d
            H[k, k] = d
            H[k + 1, k] = 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            iters += 1
            k_done = k + 1
            rel = abs(g[k + 1]) / imp_scale
            history.append(rel)
            if rel <= tol:
                stop = True
                break
        if k_done > 0:
            y = np.linalg.solve(np.triu(H[:k_done, :k_done]), g[:k_done])
            upd = V[:k_done].T @ y
            x = x + (apply_M(upd) if right else upd)
        rt = b - apply_A(x)
        r = prec_res(rt)
        if history[-1] <= tol:
            if explicit_test:
                if np.linalg.norm(rt) / exp_scale <= tol:
                    converged = True
                    break
            else:
                converged = True
                break
        if iters >= max_iters:
            break
    return x, iters, converged, history


def cg(apply_A, b, x0, apply_M=None, tol=1e-8, max_iters=1000):
    """Belos BlockCG, block size 1: PCG with test ||r||/||r0|| <= tol."""
    x = np.array(x0, dtype=np.float64, copy=True)
    r = b - apply_A(x)
    r0 = np.linalg.norm(r)
    if r0 == 0:
        return x, 0, True, [0.0]
    z = apply_M(r) if apply_M is not None else r
    p = z.copy()
    rz = r @ z
    hist = [1.0]
    it = 0
    conv = False
    while it < max_iters:
        Ap = apply_A(p)
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        it += 1
        rel = np.linalg.norm(r) / r0
        hist.append(rel)
        if rel <= tol:
            conv = True
            break
        z = apply_M(r) if apply_M is not None else r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it, conv, hist
