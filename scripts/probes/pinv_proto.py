"""Numpy prototype of the partition-inverse solve of the block-tridiagonal reduced KKT system (qp_cta_kernel.cuh,
pinv_factor / admm_block_pinv): blocks 3, 7, 11, ... are separators, the runs of <= 3 blocks between them partitions.
  factor:  PI_p = inv(A_pp);  W_p = PI_p C_p (columns: left separator, right separator);  S = A_ss - C' PI C;  Sinv;
           Z = [-Sinv W' | Sinv]  (the separator rows of the inverse of the whole matrix)
  solve:   step A: y_p = PI_p b_p  and  x_s = Z b     (independent of each other)
           step B: x_p = y_p - W_p [x_left; x_right]
Checks the formulas (index conventions of SA / SLM as assemble_factor leaves them) against a dense solve."""
import numpy as np


def build(M, NB, rng):
    N = M * NB
    G = rng.standard_normal((N, N))
    K = np.zeros((N, N))
    for p in range(M):
        a = slice(p * NB, (p + 1) * NB)
        B = rng.standard_normal((NB, NB))
        K[a, a] = B @ B.T + NB * 4 * np.eye(NB)
        if p:
            b = slice((p - 1) * NB, p * NB)
            L = rng.standard_normal((NB, NB))
            K[a, b] = L
            K[b, a] = L.T
    SA = np.stack([K[p * NB:(p + 1) * NB, p * NB:(p + 1) * NB] for p in range(M)])
    SLM = np.stack([K[p * NB:(p + 1) * NB, (p - 1) * NB:p * NB] if p else np.zeros((NB, NB)) for p in range(M)])  # K(block p, block p-1)
    return K, SA, SLM


def pinv_factor(SA, SLM, M, NB):
    seps = [b for b in range(M) if b % 4 == 3]
    parts = [list(range(4 * p, min(4 * p + 3, M))) for p in range((M + 3) // 4)]
    parts = [pb for pb in parts if pb]
    PI, W = [], []
    for pb in parts:
        n = len(pb) * NB
        A = np.zeros((n, n))
        for k, b in enumerate(pb):
            A[k * NB:(k + 1) * NB, k * NB:(k + 1) * NB] = SA[b]
            if k:
                A[k * NB:(k + 1) * NB, (k - 1) * NB:k * NB] = SLM[b]
                A[(k - 1) * NB:k * NB, k * NB:(k + 1) * NB] = SLM[b].T
        pi = np.linalg.inv(A)
        w = np.zeros((n, 2 * NB))
        fb, lb = pb[0], pb[-1]
        if fb > 0:  # left separator = block fb - 1; C_left = K(first block rows, separator cols) = SLM[fb]
            w[:, :NB] = pi[:, :NB] @ SLM[fb]
        if lb + 1 < M:  # right separator = block lb + 1; C_right = K(last block rows, sep cols) = SLM[lb + 1]'
            w[:, NB:] = pi[:, (len(pb) - 1) * NB:] @ SLM[lb + 1].T
        PI.append(pi)
        W.append(w)
    ns = len(seps)
    S = np.zeros((ns * NB, ns * NB))
    for s, b in enumerate(seps):
        S[s * NB:(s + 1) * NB, s * NB:(s + 1) * NB] = SA[b]
    for p, pb in enumerate(parts):
        fb, lb, n = pb[0], pb[-1], len(pb) * NB
        sl = p - 1 if fb > 0 else None            # index of the left separator
        sr = p if lb + 1 < M else None            # index of the right separator
        # C' W restricted to the separators adjacent to this partition
        if sl is not None:
            Cl = np.zeros((n, NB)); Cl[:NB] = SLM[fb]
            S[sl * NB:(sl + 1) * NB, sl * NB:(sl + 1) * NB] -= Cl.T @ W[p][:, :NB]
        if sr is not None:
            Cr = np.zeros((n, NB)); Cr[n - NB:] = SLM[lb + 1].T
            S[sr * NB:(sr + 1) * NB, sr * NB:(sr + 1) * NB] -= Cr.T @ W[p][:, NB:]
        if sl is not None and sr is not None:
            S[sl * NB:(sl + 1) * NB, sr * NB:(sr + 1) * NB] -= Cl.T @ W[p][:, NB:]
            S[sr * NB:(sr + 1) * NB, sl * NB:(sl + 1) * NB] -= Cr.T @ W[p][:, :NB]
    Sinv = np.linalg.inv(S) if ns else np.zeros((0, 0))
    Z = np.zeros((ns * NB, M * NB))
    for s, b in enumerate(seps):
        Z[:, b * NB:(b + 1) * NB] = Sinv[:, s * NB:(s + 1) * NB]
    for p, pb in enumerate(parts):
        fb, lb = pb[0], pb[-1]
        cols = slice(fb * NB, (lb + 1) * NB)
        if fb > 0:
            Z[:, cols] -= Sinv[:, (p - 1) * NB:p * NB] @ W[p][:, :NB].T
        if lb + 1 < M:
            Z[:, cols] -= Sinv[:, p * NB:(p + 1) * NB] @ W[p][:, NB:].T
    return parts, seps, PI, W, Z


def pinv_solve(parts, seps, PI, W, Z, b, M, NB):
    x = np.zeros(M * NB)
    xs = Z @ b                                     # step A (separator rows)
    for s, blk in enumerate(seps):
        x[blk * NB:(blk + 1) * NB] = xs[s * NB:(s + 1) * NB]
    for p, pb in enumerate(parts):
        fb, lb = pb[0], pb[-1]
        y = PI[p] @ b[fb * NB:(lb + 1) * NB]       # step A (partition rows)
        xl = x[(fb - 1) * NB:fb * NB] if fb > 0 else np.zeros(NB)
        xr = x[(lb + 1) * NB:(lb + 2) * NB] if lb + 1 < M else np.zeros(NB)
        x[fb * NB:(lb + 1) * NB] = y - W[p] @ np.concatenate([xl, xr])   # step B
    return x


def gj_rows(A):
    """The elimination as gj_rows (qp_cta_kernel.cuh) runs it: columns stay put, the thread's entry of the pivot column comes
    from the pivot row by (anti)symmetry, finished columns and pivot rows keep their stored values and carry scales applied
    at the end, the diagonal entries live in registers of their own."""
    n = A.shape[0]
    a, diag, cs, rs, inv_rs = A.copy(), np.diag(A).copy(), np.ones(n), np.ones(n), np.ones(n)
    for k in range(n):
        raw, d = a[k].copy(), diag[k]
        pv = 1.0 / d
        buf = raw.copy()
        buf[k] = 0.0
        for i in range(n):
            if i == k:
                rs[i], inv_rs[i], diag[i] = pv, d, pv
            else:
                tr = raw[i]
                f = tr if i > k else -tr * cs[i]
                g = f * pv
                diag[i] = diag[i] - g * f if i > k else diag[i] + g * f
                a[i] = a[i] - (g * inv_rs[i]) * buf
        cs[k] = -pv
    X = a * rs[:, None] * cs[None, :]
    X[np.arange(n), np.arange(n)] = diag
    return X


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for M, NB in [(15, 14), (6, 14), (5, 14), (4, 14), (8, 12), (1, 4), (3, 6), (16, 4), (7, 14)]:
        K, SA, SLM = build(M, NB, rng)
        f = pinv_factor(SA, SLM, M, NB)
        b = rng.standard_normal(M * NB)
        x = pinv_solve(*f, b, M, NB)
        print(M, NB, "max err", np.abs(x - np.linalg.solve(K, b)).max())
    for n, scale in [(42, 1.0), (42, 1e4), (42, 1e8)]:
        B = rng.standard_normal((n, n))
        D = np.diag(np.exp(rng.uniform(0, np.log(scale), n)))
        A = D @ (B @ B.T + n * np.eye(n)) @ D
        Xt = np.linalg.inv(A)
        print("gj_rows", n, scale, "rel err %.2e" % (np.abs(gj_rows(A) - Xt).max() / np.abs(Xt).max()))
