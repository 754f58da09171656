"""ORACLE (test infrastructure, not product code): sequential numpy restatement
of the reference's ADMM iteration (omgtools/problems/admm.py:584-628
ADMMProblem.dual_update with per-updater update_x 383-398, communicate 468-475,
update_z 407-440 / construct_upd_z 117-168, update_l 442-466, get_residuals
493-508, init_step 477-491, accelerate 510-554), one agent at a time, x-update through the CPU
interior-point oracle.  The z-update solves the KKT system of the reference
(f, G, h, mu) literally instead of using the precomputed projector of the
product path.  parity unpinned (no CasADi/IPOPT here)."""
import numpy as np

from omg_tools_b200.basics.spline_extra import shiftfirstknot_T, shiftoverknot_T
from oracle import ipm_c, ipm_ref


class ADMMOracle(object):
    def __init__(self, problem):
        p = self.p = problem
        self.X = p.X.copy()
        self.x_i, self.z_i, self.l_i = p.x_i.copy(), p.z_i.copy(), p.l_i.copy()
        self.x_j, self.z_ij, self.l_ij = p.x_j.copy(), p.z_ij.copy(), p.l_ij.copy()
        self.z_ji, self.l_ji = p.z_ji.copy(), p.l_ji.copy()
        self.time_prev = 0.
        self.status = None
        self.alpha, self.c_res_p = 1., None      # Nesterov state (admm.py:510-516)

    def _solve(self, x0, par):
        tb = self.p.tb
        if ipm_c.available():
            r = ipm_c.solve_batch_full(tb, x0[None], par[None], threads=1)
            return r['x'][0], int(r['status'][0]), int(r['iters'][0])
        r = ipm_ref.solve(tb, x0, par)
        return r.x, r.status, r.iters

    def _blockdiag(self, T, nblk):
        return np.kron(np.eye(nblk), T)

    def dual_update(self, t):
        p = self.p
        N, nsh, nn, L = p.N, p.nsh, p.n_nghb, p.L
        rho = p.options['rho']
        if t > 0. and int(np.round(self.time_prev / p.knot_time, 6)) < int(np.round(t / p.knot_time, 6)):
            Ts = p.shared_shift_T() if hasattr(p, 'shared_shift_T') else shiftoverknot_T(p.basis)
            for name in ('x_i', 'z_i', 'l_i', 'x_j', 'z_ij', 'l_ij', 'z_ji', 'l_ji'):
                a = getattr(self, name)
                setattr(self, name, a.reshape(-1, L).dot(Ts.T).reshape(a.shape))
            for (_, _, off, shape, T) in p.father.shifted_entries():
                for c in range(shape[1]):
                    seg = slice(off + c * shape[0], off + (c + 1) * shape[0])
                    self.X[:, seg] = self.X[:, seg].dot(np.asarray(T).T)
        self.time_prev = t
        # ---- x-update, one agent after the other -----------------------------------
        p.z_i, p.z_ji, p.l_i, p.l_ji = self.z_i, self.z_ji, self.l_i, self.l_ji
        P = p.pack_parameters(t).copy()
        self.status = np.zeros(N, dtype=int)
        self.iters = np.zeros(N, dtype=int)
        for i in range(N):
            self.X[i], self.status[i], self.iters[i] = self._solve(self.X[i], P[i])
        self.x_i = self.X[:, p.x_off:p.x_off + nsh].copy()
        # ---- communicate x --------------------------------------------------------------
        self.x_j = self.x_i[p.nghb]
        # ---- z, l, residuals ------------------------------------------------------------
        t0 = (np.round(t, 6) % p.knot_time) / p.T
        Tf, Tb = p.first_knot_transforms(t) if hasattr(p, 'shared_shift_T') else \
            shiftfirstknot_T(p.basis, t0, inverse=True)
        nblk = nsh // L * (1 + nn)
        TF, TB = self._blockdiag(Tf, nblk), self._blockdiag(Tb, nblk)
        A = p.A
        pr = dr = cr = 0.
        z_i_new, z_ij_new = np.zeros_like(self.z_i), np.zeros_like(self.z_ij)
        for i in range(N):
            x = TF.dot(np.r_[self.x_i[i], self.x_j[i].reshape(-1)])
            l = TF.dot(np.r_[self.l_i[i], self.l_ij[i].reshape(-1)])
            b = p._b_of(i)
            # admm.py:149-155
            f = -(l + rho * x)
            G = -(1. / rho) * A.dot(A.T)
            h = b + (1. / rho) * A.dot(f)
            mu = np.linalg.solve(G, h)
            z = TB.dot(-(1. / rho) * (A.T.dot(mu) + f))
            z_i_new[i], z_ij_new[i] = z[:nsh], z[nsh:].reshape(nn, nsh)
        z_i_p, z_ij_p = self.z_i, self.z_ij
        l_i_p, l_ij_p = self.l_i, self.l_ij
        self.z_i, self.z_ij = z_i_new, z_ij_new
        self.l_i = self.l_i + rho * (self.x_i - self.z_i)
        self.l_ij = self.l_ij + rho * (self.x_j - self.z_ij)
        for i in range(N):
            tf = lambda a: TF.dot(a)
            e1 = tf(np.r_[self.x_i[i] - self.z_i[i], (self.x_j[i] - self.z_ij[i]).reshape(-1)])
            e2 = tf(np.r_[self.z_i[i] - z_i_p[i], (self.z_ij[i] - z_ij_p[i]).reshape(-1)])
            pri, dri = e1.dot(e1), rho * e2.dot(e2)
            pr += pri; dr += dri; cr += rho * pri + dri
        if p.options.get('nesterov_acceleration'):
            # fast ADMM (Goldstein et al.), admm.py:510-554 with nesterov_reset
            eta = p.options.get('eta', 0.999)
            if self.c_res_p is None:
                self.c_res_p = cr / eta
            if (not p.options.get('nesterov_reset')) or cr <= eta * self.c_res_p:
                alpha_p = self.alpha
                self.alpha = 0.5 * (1. + np.sqrt(1. + 4. * alpha_p**2))
                w = (alpha_p - 1.) / self.alpha
                if not p.options.get('AMA'):       # AMA extrapolates the multipliers only
                    self.z_i = self.z_i + w * (self.z_i - z_i_p)
                    self.z_ij = self.z_ij + w * (self.z_ij - z_ij_p)
                self.l_i = self.l_i + w * (self.l_i - l_i_p)
                self.l_ij = self.l_ij + w * (self.l_ij - l_ij_p)
                self.c_res_p = cr
            else:
                self.alpha = 1.
                self.z_i, self.z_ij, self.l_i, self.l_ij = z_i_p, z_ij_p, l_i_p, l_ij_p
                self.c_res_p = self.c_res_p / eta
        # ---- communicate z, l ---------------------------------------------------------------
        for i in range(N):
            for k, j in enumerate(p.nghb[i]):
                self.z_ji[i, k] = self.z_ij[j, p.back[i, k]]
                self.l_ji[i, k] = self.l_ij[j, p.back[i, k]]
        return float(np.sqrt(pr)), float(np.sqrt(dr)), float(cr)


class DDOracle(object):
    """Sequential restatement of the reference's dual decomposition iteration
    (omgtools/problems/dualdecomposition.py: DDProblem.dual_update 279-314 = init_step,
    update_xz, communicate, update_l + get_residuals, communicate) on the problem object of
    omg_tools_b200/problems/dualdecomposition.py.  Test infrastructure."""

    def __init__(self, problem):
        p = self.p = problem
        self.X = p.X.copy()
        self.x_i, self.x_j, self.z_ij = p.x_i.copy(), p.x_j.copy(), p.z_ij.copy()
        self.l_ij, self.l_ji = p.l_ij.copy(), p.l_ji.copy()
        self.time_prev = 0.
        self.status = self.iters = None

    def dual_update(self, t):
        p = self.p
        N, nsh, nn, L = p.N, p.nsh, p.n_nghb, p.L
        rho = p.options['rho']
        if t > 0. and int(np.round(self.time_prev / p.knot_time, 6)) < int(np.round(t / p.knot_time, 6)):
            Ts = shiftoverknot_T(p.basis)                      # dualdecomposition.py:232-246
            for name in ('x_i', 'x_j', 'z_ij', 'l_ij', 'l_ji'):
                a = getattr(self, name)
                setattr(self, name, a.reshape(-1, L).dot(Ts.T).reshape(a.shape))
            for (_, _, off, shape, T) in p.father.shifted_entries():
                for c in range(shape[1]):
                    seg = slice(off + c * shape[0], off + (c + 1) * shape[0])
                    self.X[:, seg] = self.X[:, seg].dot(np.asarray(T).T)
            self.X[:, p.z_off:p.z_off + nsh * nn] = self.z_ij.reshape(N, -1)
        self.time_prev = t
        # ---- xz-update, one agent after the other (update_xz, 190-207) -------------------
        p.l_ij, p.l_ji = self.l_ij, self.l_ji
        P = p.pack_parameters(t).copy()
        self.status, self.iters = np.zeros(N, dtype=int), np.zeros(N, dtype=int)
        for i in range(N):
            r = ipm_c.solve_batch_full(p.tb, self.X[i][None], P[i][None], threads=1)
            self.X[i], self.status[i], self.iters[i] = r['x'][0], int(r['status'][0]), int(r['iters'][0])
        self.x_i = self.X[:, p.x_off:p.x_off + nsh].copy()
        self.z_ij = self.X[:, p.z_off:p.z_off + nsh * nn].reshape(N, nn, nsh).copy()
        # ---- communicate x (225-230), multiplier update (209-223), residual (248-258) ------
        self.x_j = self.x_i[p.nghb]
        self.l_ij = self.l_ij + rho * (self.x_j - self.z_ij)
        Tf, _ = p.first_knot_transforms(t)
        e = (self.x_j - self.z_ij).reshape(-1, L).dot(np.asarray(Tf).T)
        p_res = float(np.sqrt((e * e).sum()))
        # ---- communicate l: what neighbour j holds for me ------------------------------------
        self.l_ji = self.l_ij[p.nghb, p.back]
        return p_res
