"""Block-angular (doubly bordered block-diagonal) LP container and instance builders.

Host-side data contract of the KKT hot path, in the reference's block vocabulary
(PIPS-IPM/Core/Problems/DistributedProblem.cpp:2434-2553, getLocal{A,B,C,D,F,G}):

    child i :  T_i = getLocalA (my_i x n0, linking-variable columns)   W_i = getLocalB (my_i x n_i)
               C_i = getLocalC (mz_i x n0)                             D_i = getLocalD (mz_i x n_i)
               F_i = getLocalF (myl x n_i)                             G_i = getLocalG (mzl x n_i)
    root    :  A_0 (my0 x n0; the reference stores it as the root's "B",
               Readers/Distributed/DistributedTreeCallbacks.C:492-508), C_0 (mz0 x n0), F_0, G_0.

All matrices are scipy CSR with int32 indices (the reference's krowM/jcolM/M), vectors float64,
bound indicators are 0/1 float vectors (ixlow/ixupp/iclow/icupp).

Instances:
  * c1_hier_4blocks()  - literal transcription of
      Drivers/gams/simple/GAMSsmall/examples_hierarchical_approach/hier_approach_4blocks_2by2.gms:3-122
      (objective 14.0, 4 IPM iterations: Test/IntegrationTests/gamssmall_instance_data.txt:8)
  * callback_example() - literal transcription of Drivers/CallbackExample/callbackExample.cpp:32-347 (objective 14)
  * synthetic(cfg)     - the deterministic splitmix64 generator of SURVEY.md section 8(d), configs C2..C5 (+ scaled-down variants)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import scipy.sparse as sp

_U64 = np.uint64
_MASK = (1 << 64) - 1


class SplitMix64:
    """Counter-based splitmix64: element k of the stream is mix(seed + (k+1)*gamma), so whole
    arrays are produced vectorised and the same stream is trivially reproducible in C."""

    GAMMA = 0x9E3779B97F4A7C15

    def __init__(self, seed: int):
        self.state = seed & _MASK

    def u64(self, n: int) -> np.ndarray:
        k = np.arange(1, n + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            z = _U64(self.state) + k * _U64(self.GAMMA)
            self.state = int(z[-1]) if n > 0 else self.state
            z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
            z = z ^ (z >> _U64(31))
        return z

    def skip(self, n: int) -> None:
        """advance the stream by n draws without generating them"""
        self.state = (self.state + n * self.GAMMA) & _MASK

    def uniform(self, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
        u = (self.u64(n) >> _U64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        return lo + (hi - lo) * u

    def integers(self, n: int, hi: int) -> np.ndarray:
        """uniform ints in [0, hi)"""
        return (self.u64(n) % _U64(max(hi, 1))).astype(np.int64)


def _csr(m, shape) -> sp.csr_matrix:
    m = sp.csr_matrix(m, shape=shape, dtype=np.float64)
    m.sum_duplicates()
    m.sort_indices()
    m.indices = m.indices.astype(np.int32)
    m.indptr = m.indptr.astype(np.int32)
    return m


def _empty(r, c) -> sp.csr_matrix:
    return _csr(sp.csr_matrix((r, c)), (r, c))


@dataclass
class Node:
    """One node of the two-level block tree (root or scenario block)."""
    n: int
    my: int
    mz: int
    c: np.ndarray
    xlow: np.ndarray
    ixlow: np.ndarray
    xupp: np.ndarray
    ixupp: np.ndarray
    A: sp.csr_matrix          # my x n0 : root A_0 / child T_i
    b: np.ndarray
    C: sp.csr_matrix          # mz x n0 : root C_0 / child C_i (linking-variable columns)
    clow: np.ndarray
    iclow: np.ndarray
    cupp: np.ndarray
    icupp: np.ndarray
    F: sp.csr_matrix          # myl x n
    G: sp.csr_matrix          # mzl x n
    W: Optional[sp.csr_matrix] = None   # my x n  (children only)
    D: Optional[sp.csr_matrix] = None   # mz x n  (children only)


@dataclass
class BlockAngularLP:
    root: Node
    children: List[Node]
    myl: int
    mzl: int
    bl: np.ndarray
    dllow: np.ndarray
    idllow: np.ndarray
    dlupp: np.ndarray
    idlupp: np.ndarray
    name: str = "lp"
    meta: dict = field(default_factory=dict)

    # ---- sizes -------------------------------------------------------------------------
    @property
    def n0(self):
        return self.root.n

    @property
    def n_S(self):
        """Schur complement dimension n0+my0+myl+mzl (LinearSystems/DistributedRootLinearSystem.C:162)."""
        return self.root.n + self.root.my + self.myl + self.mzl

    def n_K(self, i):
        ch = self.children[i]
        return ch.n + ch.my + ch.mz

    # ---- global (flattened) view used by the host-side IPM and the full-KKT matvec --------
    def x_offsets(self):
        sizes = [self.root.n] + [c.n for c in self.children]
        return np.concatenate([[0], np.cumsum(sizes)])

    def y_offsets(self):
        """equality duals: [y0; y_1..y_N; yl]"""
        sizes = [self.root.my] + [c.my for c in self.children] + [self.myl]
        return np.concatenate([[0], np.cumsum(sizes)])

    def z_offsets(self):
        sizes = [self.root.mz] + [c.mz for c in self.children] + [self.mzl]
        return np.concatenate([[0], np.cumsum(sizes)])

    def global_matrices(self):
        """(A, C) over the flattened variable / row orderings above."""
        N = len(self.children)
        rowsA = [[self.root.A] + [None] * N]
        rowsC = [[self.root.C] + [None] * N]
        for i, ch in enumerate(self.children):
            ra = [ch.A] + [None] * N
            ra[1 + i] = ch.W
            rc = [ch.C] + [None] * N
            rc[1 + i] = ch.D
            rowsA.append(ra)
            rowsC.append(rc)
        rowsA.append([self.root.F] + [ch.F for ch in self.children])
        rowsC.append([self.root.G] + [ch.G for ch in self.children])

        def fix(rows, sizes_r):
            sizes_c = [self.root.n] + [c.n for c in self.children]
            for r, row in enumerate(rows):
                for c in range(len(row)):
                    if row[c] is None:
                        row[c] = sp.csr_matrix((sizes_r[r], sizes_c[c]))
            return sp.bmat(rows, format="csr")

        A = fix(rowsA, [self.root.my] + [c.my for c in self.children] + [self.myl])
        C = fix(rowsC, [self.root.mz] + [c.mz for c in self.children] + [self.mzl])
        return A.tocsr(), C.tocsr()

    def global_vectors(self):
        nodes = [self.root] + self.children
        cat = lambda name: np.concatenate([getattr(nd, name) for nd in nodes])
        g = dict(c=cat("c"), xlow=cat("xlow"), ixlow=cat("ixlow"), xupp=cat("xupp"), ixupp=cat("ixupp"))
        g["b"] = np.concatenate([nd.b for nd in nodes] + [self.bl])
        for nm, lk in (("clow", self.dllow), ("iclow", self.idllow), ("cupp", self.dlupp), ("icupp", self.idlupp)):
            g[nm] = np.concatenate([getattr(nd, nm) for nd in nodes] + [lk])
        return g

    def datanorm(self):
        """max |entry| over all problem data (Problems/Problem.cpp:60-84)."""
        A, C = self.global_matrices()
        g = self.global_vectors()
        vals = [np.abs(A.data).max(initial=0.0), np.abs(C.data).max(initial=0.0), np.abs(g["c"]).max(initial=0.0),
                np.abs(g["b"]).max(initial=0.0),
                np.abs(g["xlow"] * g["ixlow"]).max(initial=0.0), np.abs(g["xupp"] * g["ixupp"]).max(initial=0.0),
                np.abs(g["clow"] * g["iclow"]).max(initial=0.0), np.abs(g["cupp"] * g["icupp"]).max(initial=0.0)]
        return float(max(vals))


# ------------------------------------------------------------------------------------------
# literal instances
# ------------------------------------------------------------------------------------------
def _node_from_dense(n0, cols, Arows, brows, Crows, cupp_rows, Alink, Clink, cost, free=True, is_root=False):
    """helper for c1: build a node from dense row slices; columns `cols` are the node's own variables."""
    n = len(cols)
    my, mz = Arows.shape[0], Crows.shape[0]
    z = np.zeros
    nd = Node(n=n, my=my, mz=mz, c=np.asarray(cost, float),
              xlow=z(n), ixlow=z(n), xupp=z(n), ixupp=z(n),
              A=_csr(Arows[:, :n0], (my, n0)), b=np.asarray(brows, float),
              C=_csr(Crows[:, :n0], (mz, n0)),
              clow=z(mz), iclow=z(mz), cupp=np.asarray(cupp_rows, float), icupp=np.ones(mz),
              F=_csr(Alink[:, cols], (Alink.shape[0], n)), G=_csr(Clink[:, cols], (Clink.shape[0], n)))
    if not is_root:
        nd.W = _csr(Arows[:, cols], (my, n))
        nd.D = _csr(Crows[:, cols], (mz, n))
    return nd


def c1_hier_4blocks() -> BlockAngularLP:
    """hier_approach_4blocks_2by2.gms:3-122 : 14 free variables, 15 equalities, 15 <=-inequalities;
    stage annotations: x j1..j6 root (n0=6), j7..j14 two per block; rows i1,i2 root, i3..i10 two per block,
    i11..i15 linking (myl=mzl=5). Expected objective 14.0."""
    A = np.zeros((15, 14))
    C = np.zeros((15, 14))
    rowsA = {
        1: {1: 1, 2: 1, 3: 1, 5: 1, 6: 1}, 2: {2: 1, 4: 1, 5: -1, 6: 1},
        3: {1: 1, 2: -1, 5: 1, 7: 1, 8: 1}, 4: {2: 1, 5: -1, 7: -1, 8: 1},
        5: {1: 1, 2: -1, 6: 1, 9: 1, 10: 1}, 6: {2: 1, 6: -1, 9: -1, 10: 1},
        7: {1: 1, 2: -1, 5: 1, 11: 1, 12: 1}, 8: {2: 1, 5: -1, 11: -1, 12: 1},
        9: {1: 1, 2: -1, 6: 1, 13: 1, 14: 1}, 10: {2: 1, 6: -1, 13: -1, 14: 1},
        11: {1: 1, 2: 1, 3: 1, 5: 1, 6: 1, 8: 1, 9: 1},
        12: {1: -1, 4: 1, 5: -1, 6: 1, 10: 1, 11: 1},
        13: {1: -1, 2: -1, 3: 1, 4: 1, 13: 1, 14: 1},
        14: {1: 1, 2: -1, 5: 1, 6: -1, 7: 1, 8: 1, 9: 1, 10: 1, 11: 1, 12: 1, 13: 1, 14: 1},
        15: {1: 1, 2: 1, 5: 1, 6: 1, 8: 2, 10: 2, 12: 2, 14: 2},
    }
    rowsC = {
        1: {1: 2, 2: 2, 3: 2, 5: 2, 6: 2}, 2: {2: 2, 4: 2, 5: -2, 6: 2},
        3: {1: 2, 2: -2, 5: 2, 7: 2, 8: 2}, 4: {2: 2, 5: -2, 7: -2, 8: 2},
        5: {1: 2, 2: -2, 6: 2, 9: 2, 10: 2}, 6: {2: 2, 6: -2, 9: -2, 10: 2},
        7: {1: 2, 2: -2, 5: 2, 11: 2, 12: 2}, 8: {2: 2, 5: -2, 11: -2, 12: 2},
        9: {1: 2, 2: -2, 6: 2, 13: 2, 14: 2}, 10: {2: 2, 6: -2, 13: -2, 14: 2},
        11: {1: 2, 2: 2, 3: 2, 5: 2, 6: 2, 8: 2, 9: 2},
        12: {1: -2, 4: 2, 5: -2, 6: 2, 10: 2, 11: 2},
        13: {1: 2, 2: -2, 3: 2, 5: 2, 6: 2, 13: 2, 14: 2},
        14: {1: 2, 2: -2, 5: 2, 6: -2, 7: 2, 8: 2, 9: 2, 10: 2, 11: 2, 12: 2, 13: 2, 14: 2},
        15: {1: 2, 2: 2, 5: 2, 6: 2, 8: 4, 10: 4, 12: 4, 14: 4},
    }
    for i, row in rowsA.items():
        for j, v in row.items():
            A[i - 1, j - 1] = v
    for i, row in rowsC.items():
        for j, v in row.items():
            C[i - 1, j - 1] = v
    bA = np.array([5, 2, 3, 0, 3, 0, 3, 0, 3, 0, 7, 2, 2, 8, 12], float)
    cupp = np.array([11, 5, 7, 0.5, 7, 0.5, 7, 0.5, 6.5, 0.3, 14.5, 5, 10.1, 16, 26], float)
    n0 = 6
    link = slice(10, 15)
    root = _node_from_dense(n0, list(range(6)), A[0:2], bA[0:2], C[0:2], cupp[0:2], A[link], C[link], np.ones(6), is_root=True)
    children = []
    for k in range(4):
        cols = [6 + 2 * k, 7 + 2 * k]
        rs = slice(2 + 2 * k, 4 + 2 * k)
        children.append(_node_from_dense(n0, cols, A[rs], bA[rs], C[rs], cupp[rs], A[link], C[link], np.ones(2)))
    return BlockAngularLP(root=root, children=children, myl=5, mzl=5, bl=bA[link].copy(),
                          dllow=np.zeros(5), idllow=np.zeros(5), dlupp=cupp[link].copy(), idlupp=np.ones(5),
                          name="hier_approach_4blocks_2by2", meta=dict(expected_objective=14.0, expected_iterations=4))


def callback_example() -> BlockAngularLP:
    """Drivers/CallbackExample/callbackExample.cpp:32-347: root (n=2) + 2 scenarios (n=2, n=4), my=2, mz=1,
    myl=2, mzl=1; x >= 0, c = 2; expected objective 14 (Test/pipsipmLinkConsTest.sh:25)."""
    n0 = 2

    def mk(idx):
        n = 4 if idx == 2 else 2
        Adiag = {0: (2.0, 7.0), 1: (2.0, 5.0), 2: (2.0, 4.0)}[idx]
        A = _csr(sp.csr_matrix(np.diag(Adiag)), (2, n0))
        b = {0: (2.0, 7.0), 1: (3.0, 7.0), 2: (2.0, 7.0)}[idx]
        C = _csr(sp.csr_matrix(([2.0], ([0], [0])), shape=(1, n0)), (1, n0))
        if idx == 2:
            F = _csr(sp.csr_matrix(([1.0, 1.0, 1.0, 1.0], ([0, 0, 1, 1], [1, 3, 2, 3])), shape=(2, n)), (2, n))
        else:
            F = _csr(sp.csr_matrix(([1.0, 1.0, 1.0], ([0, 0, 1], [0, 1, 0])), shape=(2, n)), (2, n))
        G = _csr(sp.csr_matrix(([1.0], ([0], [0])), shape=(1, n)), (1, n))
        nd = Node(n=n, my=2, mz=1, c=np.full(n, 2.0), xlow=np.zeros(n), ixlow=np.ones(n), xupp=np.zeros(n),
                  ixupp=np.zeros(n), A=A, b=np.array(b), C=C, clow=np.zeros(1), iclow=np.zeros(1),
                  cupp=np.full(1, 5.0), icupp=np.ones(1), F=F, G=G)
        if idx > 0:
            if idx == 1:
                W = sp.csr_matrix(([1.0, 2.0], ([0, 1], [0, 1])), shape=(2, n))
            else:
                W = sp.csr_matrix(([0.0, 3.0], ([0, 1], [0, 1])), shape=(2, n))
            W = sp.csr_matrix(W)
            nd.W = sp.csr_matrix((W.data, W.indices.astype(np.int32), W.indptr.astype(np.int32)), shape=(2, n))
            nd.D = _csr(sp.csr_matrix(([3.0], ([0], [0])), shape=(1, n)), (1, n))
        return nd

    return BlockAngularLP(root=mk(0), children=[mk(1), mk(2)], myl=2, mzl=1, bl=np.array([6.0, 4.0]),
                          dllow=np.zeros(1), idllow=np.zeros(1), dlupp=np.array([4.0]), idlupp=np.ones(1),
                          name="callbackExample", meta=dict(expected_objective=14.0))


# ------------------------------------------------------------------------------------------
# synthetic configs (table of C2..C5 and generator 8(d))
# ------------------------------------------------------------------------------------------
CONFIGS = {
    #        N     my_i   mz_i   n_i      n0   my0  mz0  myl   mzl
    "C2": (64, 12000, 8000, 24000, 500, 100, 100, 120, 80),
    "C3": (512, 30000, 20000, 60000, 2000, 200, 200, 300, 200),
    "C4": (128, 6000, 4000, 12000, 3000, 300, 300, 2000, 1000),
    "C5": (8192, 1200, 800, 2400, 300, 50, 50, 60, 40),
    # scaled-down shapes for parity tests the CPU oracle finishes in seconds (same generator, same structure)
    "T1": (4, 60, 40, 120, 10, 4, 3, 5, 4),
    "T2": (6, 300, 200, 600, 24, 6, 5, 8, 6),
    "T3": (8, 1200, 800, 2400, 60, 10, 10, 12, 8),
    "T4": (16, 3000, 2000, 6000, 120, 20, 20, 30, 20),
    # the ROOT dimensions of C3 / C4 / C2 (n_S = 2700 / 6300 / 800) over small blocks: full-oracle parity of the root stage
    "T5": (4, 3000, 2000, 6000, 2000, 200, 200, 300, 200),
    "T6": (4, 1200, 800, 2400, 3000, 300, 300, 2000, 1000),
    "T7": (4, 1500, 1000, 3000, 500, 100, 100, 120, 80),
    # many blocks over a mid-sized root: with localize_links() the Schur complement is block-banded (sparse root path)
    "S1": (24, 300, 200, 600, 480, 20, 20, 120, 80),
}
CONFIG_ID = {"C2": 2, "C3": 3, "C4": 4, "C5": 5, "T1": 11, "T2": 12, "T3": 13, "T4": 14, "T5": 15, "T6": 16, "T7": 17, "S1": 21}
TECH_WINDOW = 48   # technology columns are drawn uniformly from the +-TECH_WINDOW columns around the row's time step


def _band_matrix(rng: SplitMix64, m: int, n: int) -> sp.csr_matrix:
    """'energy-system shaped' rows: time-step row r has a 3-nnz band (t-1,t,t+1) with sign pattern +,-,+ and
    values U[0.5,1.5], plus 1-2 technology columns drawn uniformly from the time-local window."""
    r = np.arange(m, dtype=np.int64)
    ctr = np.minimum(((2 * r + 1) * n) // (2 * m), n - 1)
    vals = rng.uniform(3 * m, 0.5, 1.5).reshape(m, 3) * np.array([1.0, -1.0, 1.0])
    cols = np.stack([ctr - 1, ctr, ctr + 1], axis=1)
    ntech = 1 + (rng.u64(m) & _U64(1)).astype(np.int64)                 # 1 or 2
    off = rng.integers(2 * m, 2 * TECH_WINDOW + 1).reshape(m, 2) - TECH_WINDOW
    tcols = ctr[:, None] + off
    tvals = rng.uniform(2 * m, 0.5, 1.5).reshape(m, 2)
    tvals[:, 1] *= (ntech == 2)
    cols = np.concatenate([cols, tcols], axis=1)
    vals = np.concatenate([vals, tvals], axis=1)
    ok = (cols >= 0) & (cols < n) & (vals != 0.0)
    rows = np.repeat(r[:, None], 5, axis=1)
    M = sp.coo_matrix((vals[ok], (rows[ok], cols[ok])), shape=(m, n))
    return _csr(M, (m, n))


def _link_cols_matrix(rng: SplitMix64, m: int, n0: int) -> sp.csr_matrix:
    """every row r gets one linking-variable entry at column (r mod n0), value -U[0.5,1.5]: every x0 column
    ('capacity') is nonzero in every block, in ceil(m/n0) spaced rows."""
    r = np.arange(m, dtype=np.int64)
    return _csr(sp.coo_matrix((-rng.uniform(m, 0.5, 1.5), (r, r % max(n0, 1))), shape=(m, n0)), (m, n0))


def _sparse_rows(rng: SplitMix64, m: int, n: int, per_row: int) -> sp.csr_matrix:
    per_row = min(per_row, n)
    r = np.repeat(np.arange(m, dtype=np.int64), per_row)
    c = rng.integers(m * per_row, n)
    v = rng.uniform(m * per_row, 0.5, 1.5)
    return _csr(sp.coo_matrix((v, (r, c)), shape=(m, n)), (m, n))


def synthetic(cfg: str = "T2", n_blocks: Optional[int] = None, block_range=None) -> BlockAngularLP:
    """Deterministic synthetic block-angular LP. seed = 0x5EED0000 + 1000*config_id + block_id (block 0 = root).
    `block_range=(lo,hi)` builds only the children lo..hi-1 (the blocks one rank owns); the root is always built.
    The LP is feasible and bounded by construction around x* = 1."""
    N, my, mz, n, n0, my0, mz0, myl, mzl = CONFIGS[cfg]
    if n_blocks is not None:
        N = n_blocks
    cid = CONFIG_ID[cfg]
    lo, hi = block_range if block_range is not None else (0, N)

    def bounds(rng, k):
        ixupp = (rng.uniform(k) < 0.10).astype(np.float64)
        xupp = rng.uniform(k, 5.0, 50.0) * ixupp
        return np.zeros(k), np.ones(k), xupp, ixupp

    # ---- root ------------------------------------------------------------------------------
    rng = SplitMix64(0x5EED0000 + 1000 * cid + 0)
    A0 = _sparse_rows(rng, my0, n0, 4)
    C0 = _sparse_rows(rng, mz0, n0, 4)
    F0 = _sparse_rows(rng, myl, n0, 4)
    G0 = _sparse_rows(rng, mzl, n0, 4)
    c0 = rng.uniform(n0, 1.0, 2.0)
    xl, ixl, xu, ixu = bounds(rng, n0)
    slack0 = rng.uniform(mz0, 0.5, 1.5)
    slackl = rng.uniform(mzl, 0.5, 1.5)
    one0 = np.ones(n0)
    root = Node(n=n0, my=my0, mz=mz0, c=c0, xlow=xl, ixlow=ixl, xupp=xu, ixupp=ixu, A=A0, b=A0 @ one0, C=C0,
                clow=np.zeros(mz0), iclow=np.zeros(mz0), cupp=C0 @ one0 + slack0, icupp=np.ones(mz0), F=F0, G=G0)
    bl = F0 @ one0
    dl = G0 @ one0 + slackl

    # ---- children --------------------------------------------------------------------------
    children = []
    for i in range(1, N + 1):
        rng = SplitMix64(0x5EED0000 + 1000 * cid + i)
        own = lo <= i - 1 < hi
        if own:
            W = _band_matrix(rng, my, n)
            D = _band_matrix(rng, mz, n)
            T = _link_cols_matrix(rng, my, n0)
            Cl = _link_cols_matrix(rng, mz, n0)
        else:
            rng.skip(8 * my + 8 * mz + my + mz)     # draws consumed by W, D, T, Cl
        F = _sparse_rows(rng, myl, n, 8)
        G = _sparse_rows(rng, mzl, n, 8)
        one = np.ones(n)
        bl = bl + F @ one           # linking rhs needs every block, owned or not
        dl = dl + G @ one
        if not own:
            continue
        c = rng.uniform(n, 1.0, 2.0)
        xl, ixl, xu, ixu = bounds(rng, n)
        slack = rng.uniform(mz, 0.5, 1.5)
        children.append(Node(n=n, my=my, mz=mz, c=c, xlow=xl, ixlow=ixl, xupp=xu, ixupp=ixu, A=T, W=W,
                             b=T @ one0 + W @ one, C=Cl, D=D, clow=np.zeros(mz), iclow=np.zeros(mz),
                             cupp=Cl @ one0 + D @ one + slack, icupp=np.ones(mz), F=F, G=G))
    return BlockAngularLP(root=root, children=children, myl=myl, mzl=mzl, bl=bl, dllow=np.zeros(mzl),
                          idllow=np.zeros(mzl), dlupp=dl, idlupp=np.ones(mzl), name=f"synthetic-{cfg}",
                          meta=dict(config=cfg, n_blocks_total=N, block_range=(lo, hi)))


def localize_links(lp: BlockAngularLP, halo: float = 0.5) -> BlockAngularLP:
    """Variant with LOCAL linking structure (the '2-links' situation the reference's sparse Schur complement exists for,
    LS/sLinsysRootAug.C:79-208): block g of N only touches the linking variables / linking rows of its own 1/N-th of the
    index range plus `halo` of a range on either side, so the root Schur complement becomes block-banded. Entries outside
    the window are dropped from A_i, C_i (columns) and F_i, G_i (rows); right-hand sides are NOT regenerated - the result
    is meant for linear-system level tests (factorize / solveCompressed on a given iterate), not for IPM runs."""
    import copy
    out = copy.copy(lp)
    N = lp.meta.get("n_blocks_total", len(lp.children))
    lo = lp.meta.get("block_range", (0, len(lp.children)))[0]
    n0 = lp.root.n

    def window(g, size):
        w = size / max(N, 1)
        return int(np.floor((g - halo) * w)), int(np.ceil((g + 1 + halo) * w))

    def keep_cols(M, a, b):
        M = M.tocoo()
        ok = (M.col >= a) & (M.col < b)
        return _csr(sp.coo_matrix((M.data[ok], (M.row[ok], M.col[ok])), shape=M.shape), M.shape)

    def keep_rows(M, a, b):
        M = M.tocoo()
        ok = (M.row >= a) & (M.row < b)
        return _csr(sp.coo_matrix((M.data[ok], (M.row[ok], M.col[ok])), shape=M.shape), M.shape)

    kids = []
    for i, ch in enumerate(lp.children):
        g = lo + i
        c = copy.copy(ch)
        a, b = window(g, n0)
        c.A, c.C = keep_cols(ch.A, a, b), keep_cols(ch.C, a, b)
        a, b = window(g, lp.myl)
        c.F = keep_rows(ch.F, a, b)
        a, b = window(g, lp.mzl)
        c.G = keep_rows(ch.G, a, b)
        kids.append(c)
    out.children = kids
    out.name = lp.name + "-local-links"
    out.meta = dict(lp.meta, local_links=halo)
    return out


def synthetic_iterate(lp: BlockAngularLP, cfg_id: int = 0):
    """Isolated-KKT benchmarking iterate: dd ~ 10^U[-3,3], nOmegaInv ~ -10^U[-3,3], seed + 777.
    Returns (dd, nomegainv) as lists over [root, child_1..child_N] plus the linking-row diagonal (last of nomegainv)."""
    lo = lp.meta.get("block_range", (0, len(lp.children)))[0]
    dd, nom = [], []
    for k, nd in enumerate([lp.root] + lp.children):
        bid = 0 if k == 0 else lo + k
        rng = SplitMix64(0x5EED0000 + 1000 * cfg_id + bid + 777)
        dd.append(10.0 ** rng.uniform(nd.n, -3.0, 3.0))
        nom.append(-(10.0 ** rng.uniform(nd.mz, -3.0, 3.0)))
    rng = SplitMix64(0x5EED0000 + 1000 * cfg_id + 777 + 999983)
    nom.append(-(10.0 ** rng.uniform(lp.mzl, -3.0, 3.0)))
    return dd, nom
