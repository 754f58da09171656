"""Batched receding-horizon driver: B independent copies of one Point2point
scenario advance in lock step, every MPC step is ONE batched solve on the GPU.

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


class BatchMPC(object):

    def __init__(self, problem, batch, update_time=0.1, jitter=0.0, seed=0, device_predict=True):
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
        dev = torch.device('cuda', self.solver.device)
        self.dev = dev
        rng = np.random.default_rng(seed)
        n, m = self.tb.n, self.tb.m
        # per-instance scenario data (host)
        self.state = np.repeat(np.asarray(self.vehicle.prediction['state'], float)[None], batch, 0)
        self.inp = np.repeat(np.asarray(self.vehicle.prediction['input'], float)[None], batch, 0)
        self.poseT = np.repeat(np.asarray(self.vehicle.poseT, float)[None], batch, 0)
        if jitter > 0:
            self.state[1:] += rng.uniform(-jitter, jitter, (batch - 1, 2))
            self.poseT[1:] += rng.uniform(-jitter, jitter, (batch - 1, 2))
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
        # cold start: linear interpolation per instance (holonomic.py:118-127)
        X0 = np.repeat(self.father.get_variables().cat[None], batch, 0)
        L = len(self.vehicle.basis)
        for k in range(2):
            X0[:, k * L:(k + 1) * L] = np.linspace(self.state[:, k], self.poseT[:, k], L).T
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
        # basis rows for the ideal prediction (value / derivative of the vehicle spline)
        self.time = 0.
        self.time_prev = 0.
        self.history = {'state': [self.state.copy()], 'iters': [], 'status': []}

    # ------------------------------------------------------------------
    def _pack_parameters(self, t):
        P, off, v = self.P, self.off, self.vehicle.label
        P[:, off[(v, 'state0')]:off[(v, 'state0')] + 2] = self.state
        P[:, off[(v, 'input0')]:off[(v, 'input0')] + 2] = self.inp
        P[:, off[(v, 'poseT')]:off[(v, 'poseT')] + 2] = self.poseT
        for o, d in zip(self.obstacles, self.obs):
            nd = o.n_dim
            for key in ('x', 'v', 'a'):
                P[:, off[(o.label, key)]:off[(o.label, key)] + nd] = d[key]
            if 'theta' in d:
                P[:, off[(o.label, 'theta')]] = d['theta'][:, 0]
        P[:, off[(self.problem.label, 't')]] = np.round(t, 6) % self.knot_time
        P[:, off[(self.problem.label, 'T')]] = self.T

    def _predict(self, Xh, t_rel, dt):
        """state/input of every instance at t + dt on its current spline."""
        basis = self.vehicle.basis
        tau = (t_rel + dt) / self.T
        B0 = basis.eval_basis([tau])[0]
        Bd, P1 = basis.derivative(1)
        B1 = Bd.eval_basis([tau])[0].dot(P1)
        L = len(basis)
        for k in range(2):
            c = Xh[:, k * L:(k + 1) * L]
            self.state[:, k] = c.dot(B0)
            self.inp[:, k] = c.dot(B1) / self.T
        return Xh

    def _predict_device(self, t_rel, dt):
        from ..solver.b200 import sample_batch
        basis = self.vehicle.basis
        tau = (t_rel + dt) / self.T
        B0 = basis.eval_basis([tau])
        Bd, P1 = basis.derivative(1)
        B1 = Bd.eval_basis([tau]).dot(P1) / self.T
        out = sample_batch(self.X, [(0, len(basis), 2, np.vstack([B0, B1]))]).cpu().numpy()
        # layout: [column][sample] with samples (value, derivative)
        self.state[:, 0], self.inp[:, 0] = out[:, 0], out[:, 1]
        self.state[:, 1], self.inp[:, 1] = out[:, 2], out[:, 3]
        return None

    def _advance_obstacles(self, dt):
        for d in self.obs:
            d['x'] = d['x'] + dt * d['v'] + 0.5 * dt * dt * d['a']
            d['v'] = d['v'] + dt * d['a']
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
        # ideal update: vehicle and obstacles move over update_time; the prediction
        # (spline value / derivative at t + update_time) is sampled on the device
        t_rel = np.round(t, 6) % self.knot_time
        Xh = self._predict_device(t_rel, self.update_time) if self.device_predict \
            else self._predict(self.X.cpu().numpy(), t_rel, self.update_time)
        self._advance_obstacles(self.update_time)
        self.history['state'].append(self.state.copy())
        self.time = np.round(t + self.update_time, 6)
        return Xh

    def run(self, n_steps):
        for _ in range(n_steps):
            self.step()
        return self.history
