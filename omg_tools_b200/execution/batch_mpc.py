"""Batched receding-horizon driver: B independent copies of one Point2point
scenario (Holonomic, Holonomic3D or Quadrotor3D vehicle) advance in lock step,
every MPC step is ONE batched solve on the GPU.

It is the batched counterpart of the reference's ``Simulator.run()`` /
``Deployer.update()`` loop (omgtools/execution/simulator.py:39-99,
deployer.py:43-79) restricted to the ideal, noise-free case (the vehicle follows
its spline; reference options ideal_prediction / ideal_update):

    per step:   predict  -> state0/input0 = spline and derivative at t + update_time
                init_step-> knot-crossing shift T.dot(coeffs) of the seg0 variables
                            (point2point.py:187-198)          [omg_shift_batch]
                set_parameters -> t, T, state0, input0, obstacle x/v/a/theta
                            (point2point.py:174-181, obstacle.py:142-155,345-348)
                solve    -> problem(x0, p, lbg, ubg)          [omg_solve_batch]

The decision variables stay resident on the device between steps (warm start);
only the parameter rows (n_par doubles per instance) travel each step.
"""
import numpy as np


class _HolonomicAdapter(object):
    """Holonomic / Holonomic3D: state = position spline value, input = its
    derivative / T (holonomic.py:87-105, 153-159)."""

    def __init__(self, mpc, vehicle, batch, jitter, rng):
        self.mpc, self.v = mpc, vehicle
        self.nd = nd = vehicle.n_dim
        rep = lambda a: np.repeat(np.asarray(a, float)[None], batch, 0)
        self.state, self.inp = rep(vehicle.prediction['state']), rep(vehicle.prediction['input'])
        self.poseT = rep(vehicle.poseT)
        if jitter > 0:
            self.state[1:] += rng.uniform(-jitter, jitter, (batch - 1, nd))
            self.poseT[1:] += rng.uniform(-jitter, jitter, (batch - 1, nd))

    def cold_start(self, X0):
        L = len(self.v.basis)
        for k in range(self.nd):
            X0[:, k * L:(k + 1) * L] = np.linspace(self.state[:, k], self.poseT[:, k], L).T

    def pack(self, P, off):
        v, nd = self.v.label, self.nd
        P[:, off[(v, 'state0')]:off[(v, 'state0')] + nd] = self.state
        P[:, off[(v, 'input0')]:off[(v, 'input0')] + nd] = self.inp
        P[:, off[(v, 'poseT')]:off[(v, 'poseT')] + nd] = self.poseT

    def predict(self, X, t_rel, dt, T, device=True):
        from ..solver.b200 import sample_batch
        basis, nd = self.v.basis, self.nd
        tau = (t_rel + dt) / T
        B0 = basis.eval_basis([tau])
        Bd, P1 = basis.derivative(1)
        B1 = Bd.eval_basis([tau]).dot(P1) / T
        L = len(basis)
        if device:
            out = sample_batch(X, [(0, L, nd, np.vstack([B0, B1]))]).cpu().numpy()
            # layout: [column][sample] with samples (value, derivative)
            for k in range(nd):
                self.state[:, k], self.inp[:, k] = out[:, 2 * k], out[:, 2 * k + 1]
        else:
            Xh = X.cpu().numpy()
            for k in range(nd):
                c = Xh[:, k * L:(k + 1) * L]
                self.state[:, k] = c.dot(B0[0])
                self.inp[:, k] = c.dot(B1[0])

    def position(self):
        return self.state


class _Quadrotor3DAdapter(object):
    """Quadrotor3D (quadrotor3d.py): the decision splines are the flat outputs
    f~, q_phi, q_theta; position and velocity follow from the double integral of
    the accelerations re-anchored at the previous prediction (splines2signals,
    quadrotor3d.py:253-275).  The integral over one update is taken exactly by
    Gauss-Legendre quadrature per knot piece on device-sampled spline values."""

    NQ = 6      # exact for the degree-9 integrand (tau1 - s) * ddx(s)

    def __init__(self, mpc, vehicle, batch, jitter, rng):
        self.mpc, self.v = mpc, vehicle
        rep = lambda a: np.repeat(np.asarray(a, float)[None], batch, 0)
        self.state, self.inp = rep(vehicle.prediction['state']), rep(vehicle.prediction['input'])
        self.poseT = rep(vehicle.poseT)
        if jitter > 0:
            self.state[1:, :3] += rng.uniform(-jitter, jitter, (batch - 1, 3))
            self.poseT[1:, :3] += rng.uniform(-jitter, jitter, (batch - 1, 3))

    def cold_start(self, X0):
        L = len(self.v.basis)
        q0 = np.tan(self.state[:, 6:8] / 2.)
        qT = np.tan(self.poseT[:, 3:5] / 2.)
        for k in range(2):
            X0[:, (k + 1) * L:(k + 2) * L] = np.linspace(q0[:, k], qT[:, k], L).T

    def pack(self, P, off):
        v = self.v.label
        st, inp = self.state, self.inp
        qp, qt = np.tan(st[:, 6] / 2.), np.tan(st[:, 7] / 2.)
        def put(name, val):
            val = np.asarray(val, dtype=float)
            val = val[:, None] if val.ndim == 1 else val
            P[:, off[(v, name)]:off[(v, name)] + val.shape[1]] = val

        put('q_phi0', qp), put('q_theta0', qt)
        put('f_til0', inp[:, 0] / ((1 + qp**2) * (1 + qt**2)))
        put('dq_phi0', 0.5 * inp[:, 1] * (1 + qp**2))
        put('dq_theta0', 0.5 * inp[:, 2] * (1 + qt**2))
        put('pos0', st[:, :3]), put('dpos0', st[:, 3:6])
        put('posT', self.poseT[:, :3])
        put('q_phiT', np.tan(self.poseT[:, 3] / 2.)), put('q_thetaT', np.tan(self.poseT[:, 4] / 2.))

    def predict(self, X, t_rel, dt, T, device=True):
        from ..solver.b200 import sample_batch
        basis, g = self.v.basis, self.v.g
        L = len(basis)
        tau0, tau1 = t_rel / T, (t_rel + dt) / T
        # quadrature nodes on [tau0, tau1], split at the knots in between
        brk = [tau0] + [k for k in np.unique(basis.knots) if tau0 + 1e-12 < k < tau1 - 1e-12] + [tau1]
        xg, wg = np.polynomial.legendre.leggauss(self.NQ)
        nodes = np.concatenate([0.5 * (b - a) * xg + 0.5 * (a + b) for a, b in zip(brk[:-1], brk[1:])])
        wts = np.concatenate([0.5 * (b - a) * wg for a, b in zip(brk[:-1], brk[1:])])
        S0 = basis.eval_basis(np.r_[nodes, tau1])
        Bd, P1 = basis.derivative(1)
        S1 = Bd.eval_basis([tau1]).dot(P1)
        ns = len(nodes) + 1
        if device:
            out = sample_batch(X, [(0, L, 3, np.vstack([S0, S1]))]).cpu().numpy()
        else:
            Xh = X.cpu().numpy()
            out = np.concatenate([Xh[:, k * L:(k + 1) * L].dot(np.vstack([S0, S1]).T)
                                  for k in range(3)], axis=1)
        col = lambda k: out[:, k * (ns + 1):(k + 1) * (ns + 1)]
        f, qp, qt = col(0)[:, :ns - 1], col(1)[:, :ns - 1], col(2)[:, :ns - 1]
        acc = np.stack([f * (1 - qp**2) * (2 * qt), -f * (1 + qt**2) * (2 * qp),
                        f * (1 - qp**2) * (1 - qt**2) - g], axis=2)          # [B, nodes, 3]
        I1 = np.einsum('q,bqk->bk', wts, acc)
        I2 = np.einsum('q,bqk->bk', wts * (tau1 - nodes), acc)
        pos, vel = self.state[:, :3], self.state[:, 3:6]
        new_pos = pos + T * vel * (tau1 - tau0) + T * T * I2
        new_vel = vel + T * I1
        f1, qp1, qt1 = col(0)[:, ns - 1], col(1)[:, ns - 1], col(2)[:, ns - 1]
        dqp1, dqt1 = col(1)[:, ns] / T, col(2)[:, ns] / T
        self.state = np.c_[new_pos, new_vel, 2 * np.arctan2(qp1, 1), 2 * np.arctan2(qt1, 1)]
        self.inp = np.c_[f1 * (1 + qp1**2) * (1 + qt1**2), 2 * dqp1 / (1 + qp1**2),
                         2 * dqt1 / (1 + qt1**2)]

    def position(self):
        return self.state[:, :3]


def _adapter_for(vehicle):
    name = type(vehicle).__name__
    if name in ('Holonomic', 'Holonomic3D'):
        return _HolonomicAdapter
    if name == 'Quadrotor3D':
        return _Quadrotor3DAdapter
    raise NotImplementedError('BatchMPC has no batched prediction for %s' % name)


class BatchMPC(object):

    def __init__(self, problem, batch, update_time=0.1, jitter=0.0, seed=0, device_predict=True,
                 device=None):
        import torch
        self.device_predict = device_predict
        self.torch = torch
        self.problem = problem
        self.solver = problem.problem
        self.father = problem.father
        self.tb = self.father.tables
        self.B = batch
        self.update_time = update_time
        self.vehicle = problem.vehicles[0]
        self.obstacles = problem.environment.obstacles
        self.T = problem.options['horizon_time']
        self.knot_time = problem.knot_time
        dev = device if device is not None else torch.device('cuda', self.solver.device)
        self.dev = dev
        rng = np.random.default_rng(seed)
        n, m = self.tb.n, self.tb.m
        # per-instance scenario data (host); the vehicle-specific part lives in the adapter
        self.veh = _adapter_for(self.vehicle)(self, self.vehicle, batch, jitter, rng)
        self.obs = []
        for o in self.obstacles:
            d = {'x': np.repeat(o.signals['position'][:, -1][None], batch, 0).astype(float),
                 'v': np.repeat(o.signals['velocity'][:, -1][None], batch, 0).astype(float),
                 'a': np.repeat(o.signals['acceleration'][:, -1][None], batch, 0).astype(float)}
            if 'theta' in o._parameters:
                d['theta'] = np.repeat(o.signals['orientation'][:, -1][None], batch, 0).astype(float)
                d['omega'] = float(o.signals['angular_velocity'][:, -1][0])
            self.obs.append(d)
        # parameter template and entry offsets
        self.P = np.repeat(self.father.set_parameters(0.).cat[None], batch, 0)
        ent = self.father._par_struct.entries
        self.off = {key: ent[key][0] for key in ent}
        # cold start per instance (holonomic.py:118-127, quadrotor3d.py:189-201)
        X0 = np.repeat(self.father.get_variables().cat[None], batch, 0)
        self.veh.cold_start(X0)
        self.X = torch.tensor(X0, device=dev)
        self.Xn = torch.empty_like(self.X)
        self.LAM = torch.empty((batch, m), dtype=torch.float64, device=dev)
        self.F = torch.empty(batch, dtype=torch.float64, device=dev)
        self.ST = torch.empty(batch, dtype=torch.int32, device=dev)
        self.IT = torch.empty(batch, dtype=torch.int32, device=dev)
        self.LB = torch.tensor(self.tb.lbg, device=dev)
        self.UB = torch.tensor(self.tb.ubg, device=dev)
        self.Pd = torch.empty((batch, self.tb.n_par), dtype=torch.float64, device=dev)
        self.blocks = [(off, shape[0], shape[1], T) for (_, _, off, shape, T)
                       in self.father.shifted_entries()]
        self.time = 0.
        self.time_prev = 0.
        self.history = {'state': [self.state.copy()], 'iters': [], 'status': []}

    # state / input / target of every instance (owned by the vehicle adapter)
    state = property(lambda self: self.veh.state)
    inp = property(lambda self: self.veh.inp)
    poseT = property(lambda self: self.veh.poseT)

    # ------------------------------------------------------------------
    def _pack_parameters(self, t):
        P, off = self.P, self.off
        self.veh.pack(P, off)
        for o, d in zip(self.obstacles, self.obs):
            nd = o.n_dim
            for key in ('x', 'v', 'a'):
                P[:, off[(o.label, key)]:off[(o.label, key)] + nd] = d[key]
            if 'theta' in d:
                P[:, off[(o.label, 'theta')]] = d['theta'][:, 0]
        P[:, off[(self.problem.label, 't')]] = np.round(t, 6) % self.knot_time
        P[:, off[(self.problem.label, 'T')]] = self.T

    def _advance_obstacles(self, dt, sample_time=0.01):
        """Obstacle motion over one update, sample by sample as the reference's
        simulator does it (obstacle.py:229-247): constant-acceleration steps plus the
        increments of the obstacle's 'trajectories' at their switching times."""
        n_samp = int(np.round(dt / sample_time, 6))
        for o, d in zip(self.obstacles, self.obs):
            inc = getattr(o, '_increments', [])
            for _ in range(n_samp):
                t0 = d.setdefault('time', 0.)
                t1 = t0 + sample_time
                d['x'] = d['x'] + sample_time * d['v'] + 0.5 * sample_time**2 * d['a']
                d['v'] = d['v'] + sample_time * d['a']
                for tm, l, val in inc:
                    if t0 + 1e-9 < tm <= t1 + 1e-9:      # (sample times accumulate rounding)
                        key = ('x', 'v', 'a')[l]
                        d[key] = d[key] + val
                d['time'] = t1
            if 'theta' in d:
                d['theta'] = d['theta'] + dt * d['omega']

    # ------------------------------------------------------------------
    def step(self):
        torch = self.torch
        t = self.time
        # knot crossing -> shift the warm start on the device
        if int(np.round(self.time_prev / self.knot_time, 6)) < int(np.round(t / self.knot_time, 6)):
            self.solver.shift_batch_device(self.X, self.blocks)
        self.time_prev = t
        self._pack_parameters(t)
        self.Pd.copy_(torch.from_numpy(self.P))
        self.solver.solve_batch_device(self.X, self.Pd, self.LB, self.UB, self.Xn,
                                       self.LAM, self.F, self.ST, self.IT)
        self.X, self.Xn = self.Xn, self.X
        self.history['iters'].append(self.IT.cpu().numpy().copy())
        self.history['status'].append(self.ST.cpu().numpy().copy())
        # ideal update: vehicle and obstacles move over update_time; the prediction at
        # t + update_time comes from spline values sampled on the device
        t_rel = np.round(t, 6) % self.knot_time
        self.veh.predict(self.X, t_rel, self.update_time, self.T, device=self.device_predict)
        self._advance_obstacles(self.update_time)
        self.history['state'].append(self.state.copy())
        self.time = np.round(t + self.update_time, 6)

    def run(self, n_steps):
        for _ in range(n_steps):
            self.step()
        return self.history
