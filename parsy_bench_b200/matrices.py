"""Synthetic SPD inputs of BASELINE.json: Dirichlet Laplacian stencils stored as
lower-half CSC, rows ascending, diagonal first — the layout ``readMatrix`` (common/Util.h:77) produces."""
import numpy as np

KINDS = {"2d5": 0, "3d7": 1, "3d27": 2}


def laplacian(kind: str, N: int):
    """Returns (n, colptr[int32 n+1], rowidx[int32 nnz], values[float64 nnz]) of tril(A)."""
    if kind not in KINDS:
        raise ValueError(f"unknown stencil {kind!r}")
    dims = 2 if kind == "2d5" else 3
    n = N ** dims
    if dims == 2:
        z, y, x = np.zeros(1, np.int64), *np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
        z = np.zeros_like(x)
    else:
        z, y, x = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    z, y, x = z.ravel(), y.ravel(), x.ravel()
    v = (z * N + y) * N + x if dims == 3 else y * N + x
    offs = []
    rz = (0, 1) if dims == 3 else (0,)
    for dz in rz:
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if (dz, dy, dx) <= (0, 0, 0):
                    continue  # only neighbours with a larger linear index (lower half)
                manh = abs(dx) + abs(dy) + abs(dz)
                if kind != "3d27" and manh != 1:
                    continue
                offs.append((dz, dy, dx))
    diag = {"2d5": 4.0, "3d7": 6.0, "3d27": 26.0}[kind]
    cols = [v]
    rows = [v]
    vals = [np.full(n, diag)]
    NZ = N if dims == 3 else 1
    for dz, dy, dx in offs:
        ok = (z + dz >= 0) & (z + dz < NZ) & (y + dy >= 0) & (y + dy < N) & (x + dx >= 0) & (x + dx < N)
        u = ((z + dz) * N + (y + dy)) * N + (x + dx) if dims == 3 else (y + dy) * N + (x + dx)
        cols.append(v[ok])
        rows.append(u[ok])
        vals.append(np.full(int(ok.sum()), -1.0))
    cols = np.concatenate(cols)
    rows = np.concatenate(rows)
    vals = np.concatenate(vals)
    order = np.lexsort((rows, cols))
    cols, rows, vals = cols[order], rows[order], vals[order]
    colptr = np.zeros(n + 1, np.int64)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr)
    return n, colptr.astype(np.int32), rows.astype(np.int32), vals.astype(np.float64)


def expected_nnz(kind: str, N: int) -> int:
    n = N * N if kind == "2d5" else N ** 3
    if kind == "2d5":
        return n + 2 * N * (N - 1)
    if kind == "3d7":
        return n + 3 * N * N * (N - 1)
    return n + 3 * N * N * (N - 1) + 6 * N * (N - 1) ** 2 + 4 * (N - 1) ** 3


def write_mtx(path, n, Ap, Ai, Ax, symmetric=True, comment=None):
    """Writes a CSC matrix column by column as a Matrix-Market coordinate file (1-based ``row col value`` triplets,
    17 significant digits) — the layout ``readMatrix`` (common/Util.h:77) expects for the lower half."""
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate real {'symmetric' if symmetric else 'general'}\n")
        if comment:
            f.write(f"% {comment}\n")
        f.write(f"{n} {n} {len(Ai)}\n")
        for j in range(n):
            for k in range(int(Ap[j]), int(Ap[j + 1])):
                f.write(f"{int(Ai[k]) + 1} {j + 1} {float(Ax[k]):.17g}\n")
