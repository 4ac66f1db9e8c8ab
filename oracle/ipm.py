"""CPU ORACLE (test infrastructure only): host-side interior-point driver restated from the reference, used to
obtain iteration-count / objective parity for a KKT backend (the CPU oracle of oracle/linsys.py or the CUDA
product through its C ABI). Nothing here is GPU work in the reference either.

Follows, on flattened global vectors:
  * outer loop / termination     InteriorPointMethod/PIPSIPMppSolver.cpp:29-83, 143-194
  * start point                  InteriorPointMethod/Solver.cpp:19-31, Variables.C:310-403
  * predictor / corrector        InteriorPointMethod/InteriorPointMethod.cpp:68-152 (PRIMAL variant)
  * Gondzio correctors           InteriorPointMethod/InteriorPointMethod.cpp:355-444, 446-456, 459-489
  * Mehrotra step length         InteriorPointMethod/InteriorPointMethod.cpp:696-744
  * residuals                    KKTFormulation/Residuals/Residuals.cpp:58-171, 220-291
  * LinearSystem::factorize      KKTFormulation/LinearSystems/LinearSystem.C:171-202, 262-294 (computeDiagonals)
  * LinearSystem::solve/XYZS     KKTFormulation/LinearSystems/LinearSystem.C:327-548
  * outer BiCGStab, system_mult  KKTFormulation/LinearSystems/LinearSystem.C:550-798, 809-844
Not restated (not reachable on the golden instances): numerical-trouble probing / small correctors / pure centering
(InteriorPointMethod.cpp:134-142, 213-220) and the scalers.

A backend exposes
    factorize(dd_nodes, nomegainv_nodes)         dd per [root, child..]; nomegainv per [root, child.., link]
    solve_compressed(rhs_nodes) -> rhs_nodes     in place; rhs_nodes[0] = [x0;y0;z0;yl;zl], rhs_nodes[1+i] = [x_i;y_i;z_i]
"""
from __future__ import annotations

import numpy as np

PIPS_EPS = 1e-13       # Utilities/pipsdef.h:34-37


class OracleBackend:
    def __init__(self, lp, n_ranks=1, blocksize=20):
        from .linsys import OracleRoot, partition_children
        N = len(lp.children)
        parts = partition_children(N, n_ranks)
        self.roots = [OracleRoot(lp, owned=range(lo, hi), blocksize=blocksize, rank=r) for r, (lo, hi) in enumerate(parts)]
        self.n_ranks = n_ranks

    def factorize(self, dd, nomegainv):
        from .linsys import OracleRoot
        for r in self.roots:
            r.put_diagonals(dd, nomegainv)
            r.initialize_kkt()
            r.assemble_local_kkt()
        OracleRoot.reduce_kkt(self.roots)
        for r in self.roots:
            r.finalize_kkt_dense()
            r.factorize_kkt()

    def solve_compressed(self, rhs):
        from .linsys import solve_compressed_multi
        if self.n_ranks == 1:
            return self.roots[0].solve_compressed(rhs)
        return solve_compressed_multi(self.roots, rhs)


class LinearSystem:
    """LinearSystem.C host logic over a backend."""

    def __init__(self, lp, backend, outer_solve=2, bicg_max_iter=75, verbose=False):
        self.lp, self.backend = lp, backend
        self.A, self.C = lp.global_matrices()
        self.AT, self.CT = self.A.T.tocsr(), self.C.T.tocsr()
        g = lp.global_vectors()
        self.g = g
        self.xo, self.yo, self.zo = lp.x_offsets(), lp.y_offsets(), lp.z_offsets()
        self.outer_solve = outer_solve
        self.bicg_tol = 1e-10
        self.bicg_max_iter = bicg_max_iter
        self.bicg_eps = 1e-15
        self.verbose = verbose
        self.stats = dict(n_factorize=0, n_solve=0, n_solve_compressed=0, bicg_iters=[])
        self.bicg_converged, self.bicg_skipped, self.bicg_relres, self.bicg_niter = True, True, 0.0, 0

    # ---- node <-> flat layout ----------------------------------------------------------------
    def join_rhs(self, x, y, z):
        N = len(self.lp.children)
        xo, yo, zo = self.xo, self.yo, self.zo
        out = [np.concatenate([x[xo[0]:xo[1]], y[yo[0]:yo[1]], z[zo[0]:zo[1]], y[yo[N + 1]:yo[N + 2]], z[zo[N + 1]:zo[N + 2]]])]
        for i in range(N):
            out.append(np.concatenate([x[xo[1 + i]:xo[2 + i]], y[yo[1 + i]:yo[2 + i]], z[zo[1 + i]:zo[2 + i]]]))
        return out

    def separate_vars(self, nodes):
        N = len(self.lp.children)
        lp = self.lp
        x = np.empty(self.xo[-1]); y = np.empty(self.yo[-1]); z = np.empty(self.zo[-1])
        r = lp.root
        b0 = nodes[0]
        x[:r.n] = b0[:r.n]
        y[:r.my] = b0[r.n:r.n + r.my]
        z[:r.mz] = b0[r.n + r.my:r.n + r.my + r.mz]
        y[self.yo[N + 1]:] = b0[r.n + r.my + r.mz:r.n + r.my + r.mz + lp.myl]
        z[self.zo[N + 1]:] = b0[r.n + r.my + r.mz + lp.myl:]
        for i, ch in enumerate(lp.children):
            b = nodes[1 + i]
            x[self.xo[1 + i]:self.xo[2 + i]] = b[:ch.n]
            y[self.yo[1 + i]:self.yo[2 + i]] = b[ch.n:ch.n + ch.my]
            z[self.zo[1 + i]:self.zo[2 + i]] = b[ch.n + ch.my:]
        return x, y, z

    def split_nodes(self, v, offs, with_link):
        N = len(self.lp.children)
        out = [v[offs[k]:offs[k + 1]] for k in range(N + 1)]
        if with_link:
            out.append(v[offs[N + 1]:offs[N + 2]])
        return out

    # ---- factorize -----------------------------------------------------------------------------
    def factorize(self, it):
        g = self.g
        dd = np.zeros_like(it["x"])
        ml, mu_ = g["ixlow"] > 0, g["ixupp"] > 0
        dd[ml] += it["gamma"][ml] / it["v"][ml]
        dd[mu_] += it["phi"][mu_] / it["w"][mu_]
        om = np.zeros_like(it["z"])
        cl, cu = g["iclow"] > 0, g["icupp"] > 0
        om[cl] += it["lam"][cl] / it["t"][cl]
        om[cu] += it["pi"][cu] / it["u"][cu]
        nom = np.zeros_like(om)
        nz = om != 0
        nom[nz] = -1.0 / om[nz]                  # safe_invert + negate
        self.dd, self.nomegainv = dd, nom
        self.backend.factorize(self.split_nodes(dd, self.xo, False), self.split_nodes(nom, self.zo, True))
        self.stats["n_factorize"] += 1

    # ---- system_mult: [dd A^T C^T; A 0 0; C 0 nOmegaInv] ------------------------------------------
    def system_mult(self, x, y, z):
        rx = self.dd * x + self.AT @ y + self.CT @ z
        ry = self.A @ x
        rz = self.C @ x + self.nomegainv * z
        return rx, ry, rz

    def _precond(self, vx, vy, vz):
        nodes = self.join_rhs(vx, vy, vz)
        self.backend.solve_compressed(nodes)
        self.stats["n_solve_compressed"] += 1
        return self.separate_vars(nodes)

    # ---- solveCompressedBiCGStab (LinearSystem.C:550-798), vectors kept as (x,y,z) triples ----------
    def _bicgstab(self, bx, by, bz):
        cat = lambda t: np.concatenate(t)
        b = cat((bx, by, bz))
        nx, ny = len(bx), len(by)
        split = lambda v: (v[:nx], v[nx:nx + ny], v[nx + ny:])
        K = lambda v: cat(self.system_mult(*split(v)))
        P = lambda v: cat(self._precond(*split(v)))
        tol = self.bicg_tol
        nb = np.linalg.norm(b)
        tolb = max(nb * tol, self.bicg_eps)
        x = P(b)
        r = b - K(x)
        nr = np.linalg.norm(r)
        best_x, min_nr = x.copy(), nr
        self.bicg_niter = 0
        if nr <= tolb:
            self.bicg_skipped, self.bicg_converged = True, True
            self.bicg_relres = nr / nb if nb > 0 else 0.0
            self.stats["bicg_iters"].append(0)
            return split(x)
        self.bicg_skipped = False
        r0 = r / nr
        rho = omega = alpha = 1.0
        converged = False
        ndiv = 0
        nstag = 0
        p = v = None
        res = nr
        it_done = 0
        for it in range(self.bicg_max_iter):
            it_done = it
            rho_last = rho
            rho = r0 @ r
            if abs(rho) < 1e-300:
                break
            if it == 0:
                p = r.copy()
            else:
                beta = (rho / rho_last) * (alpha / omega)
                p = r + beta * (p - omega * v)
            dx = P(p)
            v = K(dx)
            rtv = r0 @ v
            if abs(rtv) < 1e-300:
                break
            alpha = rho / rtv
            nstag = nstag + 1 if abs(alpha) * np.linalg.norm(dx) <= self.bicg_eps * np.linalg.norm(x) else 0
            x = x + alpha * dx
            r = r - alpha * v
            nr = np.linalg.norm(r)
            if nr <= tolb:
                r = b - K(x)
                res = np.linalg.norm(r)
                if res <= tolb:
                    converged = True
                    break
                min_nr = np.linalg.norm(b - K(best_x))
                nr = res
            if nr < min_nr:
                min_nr, best_x = nr, x.copy()
            dx = P(r)
            t = K(dx)
            tt = t @ t
            if tt < 1e-300:
                break
            omega = (t @ r) / tt
            nstag = nstag + 1 if abs(omega) * np.linalg.norm(dx) <= self.bicg_eps * np.linalg.norm(x) else 0
            x = x + omega * dx
            r = r - omega * t
            nr = np.linalg.norm(r)
            if nr <= tolb or nstag >= 4:
                r = b - K(x)
                res = np.linalg.norm(r)
                if res <= tolb:
                    converged = True
                    break
                min_nr = np.linalg.norm(b - K(best_x))
                nr = res
            else:
                ndiv = ndiv + 1 if nr >= min_nr else 0
                if ndiv > 4:
                    x = best_x.copy()
                    res = np.linalg.norm(b - K(x))
                    break
            if nr < min_nr:
                min_nr, best_x = nr, x.copy()
            if nstag >= 4:
                if min_nr < nr:
                    x = best_x.copy()
                res = np.linalg.norm(b - K(x))
                break
            if abs(omega) < 1e-300:
                break
        else:
            res = np.linalg.norm(b - K(x))
        self.bicg_niter = max(it_done + 1, 1)
        self.bicg_converged = converged
        self.bicg_relres = res / nb if nb > 0 else 0.0
        self.stats["bicg_iters"].append(self.bicg_niter)
        return split(x)

    # ---- solveXYZS (LinearSystem.C:449-548) -------------------------------------------------------
    def solve_xyzs(self, stepx, stepy, stepz, steps):
        stepz = stepz - self.nomegainv * steps
        if self.outer_solve == 0:
            sx, sy, sz = self._precond(stepx, stepy, stepz)
        elif self.outer_solve == 3:
            # the backend runs solveCompressedBiCGStab itself (the CUDA product's pipsgpu_kkt_solve): the driver only hands
            # over the right-hand side in node layout and reads back status / iteration count / relative residual
            nodes = self.join_rhs(stepx, stepy, stepz)
            status, iters, relres = self.backend.solve_bicgstab(nodes, tol=self.bicg_tol, max_iter=self.bicg_max_iter)
            self.bicg_converged, self.bicg_skipped = status in (0, 1), status == 1
            self.bicg_niter, self.bicg_relres = iters, relres
            self.stats["bicg_iters"].append(iters)
            sx, sy, sz = self.separate_vars(nodes)
        else:
            sx, sy, sz = self._bicgstab(stepx, stepy, stepz)
        sy = -sy
        sz = -sz
        steps = -((steps - sz) * self.nomegainv)
        return sx, sy, sz, steps

    # ---- LinearSystem::solve (LinearSystem.C:327-447) ------------------------------------------------
    def solve(self, it, res):
        g = self.g
        ixl, ixu, icl, icu = g["ixlow"] > 0, g["ixupp"] > 0, g["iclow"] > 0, g["icupp"] > 0
        self.stats["n_solve"] += 1
        rx = res["rQ"].copy()
        gv = np.zeros_like(rx); gv[ixl] = it["gamma"][ixl] / it["v"][ixl]
        rx += gv * res["rv"]
        rx[ixl] += res["rgamma"][ixl] / it["v"][ixl]
        pw = np.zeros_like(rx); pw[ixu] = it["phi"][ixu] / it["w"][ixu]
        rx += pw * res["rw"]
        rx[ixu] -= res["rphi"][ixu] / it["w"][ixu]
        rs = res["rz"].copy()
        lt = np.zeros_like(rs); lt[icl] = it["lam"][icl] / it["t"][icl]
        rs += lt * res["rt"]
        rs[icl] += res["rlambda"][icl] / it["t"][icl]
        pu = np.zeros_like(rs); pu[icu] = it["pi"][icu] / it["u"][icu]
        rs += pu * res["ru"]
        rs[icu] -= res["rpi"][icu] / it["u"][icu]
        sx, sy, sz, ss = self.solve_xyzs(rx, res["rA"].copy(), res["rC"].copy(), rs)
        st = dict(x=sx, y=sy, z=sz, s=ss)
        z0 = np.zeros_like
        st["t"] = z0(ss); st["lam"] = z0(ss); st["u"] = z0(ss); st["pi"] = z0(ss)
        st["v"] = z0(sx); st["gamma"] = z0(sx); st["w"] = z0(sx); st["phi"] = z0(sx)
        st["t"][icl] = (ss - res["rt"])[icl]
        st["lam"][icl] = ((res["rlambda"] - it["lam"] * st["t"])[icl]) / it["t"][icl]
        st["u"][icu] = (res["ru"] - ss)[icu]
        st["pi"][icu] = ((res["rpi"] - it["pi"] * st["u"])[icu]) / it["u"][icu]
        st["v"][ixl] = (sx - res["rv"])[ixl]
        st["gamma"][ixl] = ((res["rgamma"] - it["gamma"] * st["v"])[ixl]) / it["v"][ixl]
        st["w"][ixu] = (res["rw"] - sx)[ixu]
        st["phi"][ixu] = ((res["rphi"] - it["phi"] * st["w"])[ixu]) / it["w"][ixu]
        return st


PRIMAL_KEYS = ("x", "s", "t", "u", "v", "w")
DUAL_KEYS = ("y", "z", "lam", "pi", "gamma", "phi")
PAIRS = (("t", "lam", "iclow"), ("u", "pi", "icupp"), ("v", "gamma", "ixlow"), ("w", "phi", "ixupp"))


class IPM:
    """Mehrotra predictor-corrector with Gondzio multiple centrality correctors, PRIMAL variant (one step length),
    as run by the reference's callback example and the `TestGamssmallNoSettings` configuration (no scaler, no presolve)."""

    def __init__(self, lp, backend, verbose=False, max_iterations=300, outer_solve=2, gondzio_max_correctors=3,
                 n_linesearch_points=10):
        self.lp = lp
        self.ls = LinearSystem(lp, backend, outer_solve=outer_solve, verbose=verbose)
        self.g = self.ls.g
        self.A, self.C = self.ls.A, self.ls.C
        self.verbose = verbose
        self.max_iterations = max_iterations
        self.mutol, self.artol = 1e-6, 1e-4
        self.dnorm = lp.datanorm()
        self.maximum_correctors = gondzio_max_correctors
        self.n_ls = n_linesearch_points
        self.step_factor0, self.step_factor1, self.acceptance_tolerance = 0.3, 1.5, 0.01
        self.beta_min, self.beta_max, self.tsig = 0.1, 10.0, 3.0
        self.gamma_f, self.steplength_factor = 0.99, 0.99999999
        self.gamma_a = 1.0 / (1.0 - self.gamma_f)
        g = self.g
        self.mask = dict(ixlow=g["ixlow"] > 0, ixupp=g["ixupp"] > 0, iclow=g["iclow"] > 0, icupp=g["icupp"] > 0)
        self.ncomp = int(sum(m.sum() for m in self.mask.values()))
        self.history = []

    # ---- Variables --------------------------------------------------------------------------------
    def mu(self, it):
        if self.ncomp == 0:
            return 0.0
        return sum(float(it[a] @ it[b]) for a, b, _ in PAIRS) / self.ncomp

    def mustep(self, it, st, ap, ad):
        if self.ncomp == 0:
            return 0.0
        return sum(float((it[a] + ap * st[a]) @ (it[b] + ad * st[b])) for a, b, _ in PAIRS) / self.ncomp

    def add(self, it, st, ap, ad=None):
        ad = ap if ad is None else ad
        for k in PRIMAL_KEYS:
            it[k] = it[k] + ap * st[k]
        for k in DUAL_KEYS:
            it[k] = it[k] + ad * st[k]

    def fraction_to_boundary(self, it, st, fraction=1.0):
        length = 1.0
        for a, b, m in PAIRS:
            msk = self.mask[m]
            for k in (a, b):
                v, s = it[k][msk], st[k][msk]
                sel = (v >= 0) & (s < 0)
                if sel.any():
                    length = min(length, float(np.min(-fraction * v[sel] / s[sel])))
        return length

    def find_blocking(self, it, st):
        """Variables::find_blocking (single step length) + VectorUtilities.h:12-66."""
        bound, first_or_second = 1.0, 0
        pv = ps = dv = ds = 0.0
        for a, b, m in PAIRS:
            msk = self.mask[m]
            if not msk.any():
                continue
            w, ws, u, us = it[a][msk], st[a][msk], it[b][msk], st[b][msk]
            for kind, (val, stp) in ((1, (w, ws)), (2, (u, us))):
                sel = (val > 0) & (stp < 0)
                if sel.any():
                    ratios = np.full(len(val), np.inf)
                    ratios[sel] = -val[sel] / stp[sel]
                    j = int(np.argmin(ratios))
                    if ratios[j] <= bound:
                        bound = float(ratios[j]); first_or_second = kind
                        pv, ps, dv, ds = float(w[j]), float(ws[j]), float(u[j]), float(us[j])
        return bound, first_or_second, pv, ps, dv, ds

    # ---- Residuals -----------------------------------------------------------------------------------
    def evaluate(self, it):
        g, m = self.g, self.mask
        res = {}
        rQ = g["c"].copy()
        gap = float(rQ @ it["x"])
        rQ -= self.A.T @ it["y"] + self.C.T @ it["z"]
        rQ[m["ixlow"]] -= it["gamma"][m["ixlow"]]
        rQ[m["ixupp"]] += it["phi"][m["ixupp"]]
        res["rQ"] = rQ
        res["rA"] = self.A @ it["x"] - g["b"]
        gap -= float(g["b"] @ it["y"])
        res["rC"] = self.C @ it["x"] - it["s"]
        rz = it["z"].copy()
        rz[m["iclow"]] -= it["lam"][m["iclow"]]
        rz[m["icupp"]] += it["pi"][m["icupp"]]
        res["rz"] = rz
        z0 = np.zeros_like
        res["rt"] = z0(it["s"]); res["ru"] = z0(it["s"]); res["rv"] = z0(it["x"]); res["rw"] = z0(it["x"])
        res["rt"][m["iclow"]] = (it["s"] - g["clow"] - it["t"])[m["iclow"]]
        gap -= float((g["clow"] * g["iclow"]) @ it["lam"])
        res["ru"][m["icupp"]] = (it["s"] - g["cupp"] + it["u"])[m["icupp"]]
        gap += float((g["cupp"] * g["icupp"]) @ it["pi"])
        res["rv"][m["ixlow"]] = (it["x"] - g["xlow"] - it["v"])[m["ixlow"]]
        gap -= float((g["xlow"] * g["ixlow"]) @ it["gamma"])
        res["rw"][m["ixupp"]] = (it["x"] - g["xupp"] + it["w"])[m["ixupp"]]
        gap += float((g["xupp"] * g["ixupp"]) @ it["phi"])
        res["norm"] = max(np.abs(res[k]).max(initial=0.0) for k in ("rQ", "rA", "rC", "rz", "rt", "ru", "rv", "rw"))
        res["gap"] = gap
        for k in ("rlambda", "rpi"):
            res[k] = z0(it["s"])
        for k in ("rgamma", "rphi"):
            res[k] = z0(it["x"])
        return res

    def set_complementarity(self, res, v, alpha):
        m = self.mask
        for (a, b, ind), key in zip(PAIRS, ("rlambda", "rpi", "rgamma", "rphi")):
            r = v[a] * v[b]
            r[m[ind]] += alpha
            r[~m[ind]] = 0.0
            res[key] = r

    def zero_linear(self, res):
        out = {k: np.zeros_like(v) if isinstance(v, np.ndarray) else v for k, v in res.items()}
        return out

    def project_r3(self, res, rmin, rmax):
        m = self.mask
        for (a, b, ind), key in zip(PAIRS, ("rlambda", "rpi", "rgamma", "rphi")):
            v = res[key]
            out = np.where(v < rmin, rmin - v, np.where(v > rmax, rmax - v, 0.0))
            out = np.maximum(out, -rmax)
            out[~m[ind]] = 0.0
            res[key] = out

    # ---- helpers ----------------------------------------------------------------------------------------
    def neg(self, st):
        return {k: -v for k, v in st.items()}

    def alpha_weight_candidate(self, it, pred, corr, alpha_pred):
        best_a, best_w = -1.0, -1.0
        wmin = alpha_pred * alpha_pred
        wlen = 1.0 - wmin
        for n in range(self.n_ls + 1):
            wc = min(1.0, wmin + (wlen / self.n_ls) * n)
            tmp = {k: pred[k] + wc * corr[k] for k in pred}
            a = self.fraction_to_boundary(it, tmp)
            if a > best_a:
                best_a, best_w = a, wc
        return best_a, best_w

    # ---- main loop ------------------------------------------------------------------------------------------
    def solve(self):
        g, m = self.g, self.mask
        nx, ny, nz = len(g["c"]), len(g["b"]), len(g["clow"])
        pn = np.sqrt(self.dnorm)
        it = dict(x=np.zeros(nx), s=np.zeros(nz), y=np.zeros(ny), z=np.zeros(nz))
        for a, b, ind in PAIRS:
            k = nz if a in ("t", "u") else nx
            it[a] = np.where(m[ind], pn, 0.0) * np.ones(k)
            it[b] = np.where(m[ind], pn, 0.0) * np.ones(k)
        # Solver::solve_linear_system : start point
        self.ls.bicg_tol = 1e-10
        res = self.evaluate(it)
        self.set_complementarity(res, it, 0.0)
        self.ls.factorize(it)
        step = self.neg(self.ls.solve(it, res))
        self.add(it, step, 1.0)
        viol = 0.0
        for a, b, ind in PAIRS:
            if m[ind].any():
                viol = min(viol, float(it[a][m[ind]].min()), float(it[b][m[ind]].min()))
        shift = 1e3 + 2 * (-viol)
        for a, b, ind in PAIRS:
            it[a][m[ind]] += shift
            it[b][m[ind]] += shift

        iteration = 0
        status = "MAX_ITS_EXCEEDED"
        phi_min = []
        while True:
            if iteration >= self.max_iterations - 1:
                break
            res = self.evaluate(it)
            mu = self.mu(it)
            gap, rnorm = abs(res["gap"]), res["norm"]
            phi = (rnorm + gap) / self.dnorm
            phi_min.append(phi if not phi_min else min(phi, phi_min[-1]))
            obj = float(g["c"] @ it["x"])
            self.history.append(dict(iteration=iteration, mu=mu, rnorm=rnorm, gap=gap, objective=obj))
            if self.verbose:
                print(f"it {iteration:3d} mu {mu:.3e} rnorm {rnorm:.3e} gap {gap:.3e} obj {obj:.10g}")
            if mu <= self.mutol and rnorm <= self.artol * self.dnorm:
                status = "SUCCESSFUL_TERMINATION"
                break
            if iteration >= 10 and phi >= 1e-8 and phi >= 1e4 * phi_min[-1]:
                status = "INFEASIBLE"
                break
            # ---- predictor (compute_predictor_step) ----
            self.ls.bicg_tol = 1e-8 if iteration <= 3 else (1e-9 if iteration <= 7 else 1e-10)
            self.set_complementarity(res, it, 0.0)
            self.ls.factorize(it)
            step = self.neg(self.ls.solve(it, res))
            tau = max(0.99, 1.0 - mu)
            alpha = self.fraction_to_boundary(it, step, tau)
            # ---- corrector (PrimalInteriorPointMethod::compute_corrector_step) ----
            sigma = (self.mustep(it, step, alpha, alpha) / mu) ** self.tsig
            cres = self.zero_linear(res)
            self.set_complementarity(cres, step, -sigma * mu)
            corr = self.neg(self.ls.solve(it, cres))
            alpha_cand, weight = self.alpha_weight_candidate(it, step, corr, alpha)
            step = {k: step[k] + weight * corr[k] for k in step}
            # ---- Gondzio correction loop ----
            rmin, rmax = sigma * mu * self.beta_min, sigma * mu * self.beta_max
            ncorr = 0
            while ncorr < self.maximum_correctors and alpha_cand < 1.0 - PIPS_EPS:
                alpha_target = min(self.step_factor1 * alpha_cand + self.step_factor0, 1.0)
                trial = {k: it[k] + alpha_target * step[k] for k in it}
                self.set_complementarity(cres, trial, 0.0)
                self.project_r3(cres, rmin, rmax)
                corr = self.ls.solve(it, cres)            # note: NOT negated (compute_gondzio_corrector)
                alpha_enh, wcand = self.alpha_weight_candidate(it, step, corr, alpha_target)
                if abs(alpha_enh - 1.0) <= PIPS_EPS:
                    step = {k: step[k] + wcand * corr[k] for k in step}
                    alpha_cand = alpha_enh
                    ncorr += 1
                    break
                elif alpha_enh >= (1.0 + self.acceptance_tolerance) * alpha_cand:
                    step = {k: step[k] + wcand * corr[k] for k in step}
                    alpha_cand = alpha_enh
                    ncorr += 1
                else:
                    break
            # ---- Mehrotra step length ----
            amax, fos, pv, ps, dv, ds = self.find_blocking(it, step)
            mu_full = self.mustep(it, step, amax, amax) / self.gamma_a
            if fos == 0:
                a = 1.0
            elif fos == 1:
                a = (-pv + mu_full / (dv + amax * ds)) / ps
            else:
                a = (-dv + mu_full / (pv + amax * ps)) / ds
            a = min(a, amax)
            a = max(a, self.gamma_f * amax)
            a *= self.steplength_factor
            self.add(it, step, a)
            iteration += 1
        self.iterate = it
        self.iterations = iteration
        self.status = status
        self.objective = float(g["c"] @ it["x"])
        return status, self.objective, iteration
