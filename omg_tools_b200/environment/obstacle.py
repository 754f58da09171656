"""Obstacles: position spline built from (x, v, a) parameters, hyperplane rows
on the obstacle side, rotating obstacles through a piecewise NURBS circle.

Follows the reference's ``omgtools/environment/obstacle.py``: init 80-121
(degree-2 Bezier position from pos/vel/acc shifted back by t), parameters
142-155, Obstacle2D.init 292-332 (cos/sin/weight splines), collision rows
334-343 (2D) / 528-533 (3D), theta parameter 345-348.  Simulation is the
piecewise-constant-velocity/acceleration model without bouncing (geometry and
bouncing are outside the hot path)."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/environment/obstacle.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
from __future__ import division

import numpy as np

from ..basics.optilayer import OptiChild, inf
from ..basics.poly import cos, sin
from ..basics.spline_extra import get_interval_T
from ..basics.spline import BSplineBasis, BSpline


class Obstacle(object):

    def __new__(cls, initial, shape, simulation=None, options=None):
        simulation = simulation or {}
        options = options or {}
        if shape.n_dim == 2:
            return Obstacle2D(initial, shape, simulation, options)
        if shape.n_dim == 3:
            return Obstacle3D(initial, shape, simulation, options)


class ObstaclexD(OptiChild):

    def __init__(self, initial, shape, simulation, options):
        OptiChild.__init__(self, 'obstacle')
        self.simulation = simulation
        if 'trajectories' in simulation:
            for key in simulation['trajectories']:
                if 0 in simulation['trajectories'][key]['time']:
                    initial[key] = simulation['trajectories'][key]['values'][0]
        self.set_default_options()
        self.set_options(options)
        self.shape = shape
        self.n_dim = shape.n_dim
        self.basis = BSplineBasis([0, 0, 0, 1, 1, 1], 2)
        self.initial = initial
        self.prepare_simulation(initial, simulation)

    def set_default_options(self):
        self.options = {'draw': True, 'avoid': True, 'spline_traj': False,
                        'spline_params': {'knots': [0, 0, 0, 1, 1, 1],
                                          'degree': 2, 'coeffs': [0, 0, 0]},
                        'bounce': False}

    def set_options(self, options):
        self.options.update(options)

    # ------------------------------------------------------------------
    # optimization modelling
    # ------------------------------------------------------------------

    def init(self, horizon_times=None):
        if self.options['spline_traj'] is False:
            x = self.define_parameter('x', self.n_dim)
            v = self.define_parameter('v', self.n_dim)
            a = self.define_parameter('a', self.n_dim)
            # x, v, a hold at the current time; shift back to horizon start
            self.t = self.define_symbol('t')
            if horizon_times is None:
                self.T = self.define_symbol('T')
            elif not isinstance(horizon_times, list):
                horizon_times = [horizon_times]
            v0 = v - self.t * a
            x0 = x - self.t * v0 - 0.5 * (self.t**2) * a
            a0 = a
            if horizon_times:
                pos0 = x0
                self.pos_spline = [0] * self.n_dim
                for horizon_time in horizon_times:
                    for k in range(self.n_dim):
                        self.pos_spline[k] = BSpline(self.basis, np.array([
                            pos0[k],
                            0.5 * v0[k] * horizon_time + pos0[k],
                            pos0[k] + v0[k] * horizon_time +
                            0.5 * a0[k] * (horizon_time**2)], dtype=object))
                    pos0 = [self.pos_spline[k](1)[0] for k in range(self.n_dim)]
            else:
                self.pos_spline = [BSpline(self.basis, np.array([
                    x0[k], 0.5 * v0[k] * self.T + x0[k],
                    x0[k] + v0[k] * self.T + 0.5 * a0[k] * (self.T**2)],
                    dtype=object)) for k in range(self.n_dim)]
        else:
            self.basis = BSplineBasis(self.options['spline_params']['knots'],
                                      self.options['spline_params']['degree'])
            traj_coeffs = self.define_parameter(
                'traj_coeffs', len(self.basis), self.n_dim)
            traj_coeffs = np.asarray(traj_coeffs).reshape(len(self.basis), -1)
            self.pos_spline = [BSpline(self.basis, traj_coeffs[:, k])
                               for k in range(self.n_dim)]
        checkpoints, _ = self.shape.get_checkpoints()
        self.checkpoints = np.atleast_1d(self.define_parameter(
            'checkpoints', len(checkpoints) * self.n_dim))
        self.rad = np.atleast_1d(self.define_parameter('rad', len(checkpoints)))

    def define_collision_constraints(self, hyperplanes):
        raise ValueError('Please implement this method.')

    def set_parameters(self, current_time):
        parameters = {self: {}}
        if not self.options['spline_traj']:
            parameters[self]['x'] = self.signals['position'][:, -1]
            parameters[self]['v'] = self.signals['velocity'][:, -1]
            parameters[self]['a'] = self.signals['acceleration'][:, -1]
        else:
            parameters[self]['traj_coeffs'] = self.options['spline_params']['coeffs']
        checkpoints, rad = self.shape.get_checkpoints()
        parameters[self]['checkpoints'] = np.reshape(
            checkpoints, (len(checkpoints) * self.n_dim, ))
        parameters[self]['rad'] = rad
        return parameters

    # ------------------------------------------------------------------
    # host-side state propagation
    # ------------------------------------------------------------------

    def set_state(self, dictionary):
        for key in ['position', 'velocity', 'acceleration']:
            if key in dictionary:
                self.signals[key] = np.c_[dictionary[key]]
            else:
                self.signals[key] = np.zeros((self.n_dim, 1))

    def prepare_simulation(self, initial, simulation):
        """Increments of (pos, vel, acc) at given times, reference 168-228."""
        self._increments = []      # (time, key index, values)
        if 'trajectories' in simulation:
            for key, traj in simulation['trajectories'].items():
                l = ['position', 'velocity', 'acceleration'].index(key)
                for tm, val in zip(traj['time'], traj['values']):
                    if tm != 0:
                        self._increments.append(
                            (float(tm), l, np.asarray(val, dtype=float)))
        self._increments.sort(key=lambda e: e[0])
        self.signals = {'time': np.array([0.])}
        for key in ['position', 'velocity', 'acceleration']:
            if key in initial:
                self.signals[key] = np.c_[np.asarray(initial[key], dtype=float)]
            else:
                self.signals[key] = np.zeros((self.n_dim, 1))

    def simulate(self, simulation_time, sample_time):
        n_samp = int(np.round(simulation_time / sample_time, 6))
        for _ in range(n_samp):
            t0 = self.signals['time'][-1]
            t1 = t0 + sample_time
            st = [self.signals[k][:, -1].copy()
                  for k in ('position', 'velocity', 'acceleration')]
            st[0] = st[0] + sample_time * st[1] + 0.5 * sample_time**2 * st[2]
            st[1] = st[1] + sample_time * st[2]
            for tm, l, val in self._increments:
                if t0 + 1e-9 < tm <= t1 + 1e-9:      # (sample times accumulate rounding)
                    st[l] = st[l] + val
            for k, key in enumerate(('position', 'velocity', 'acceleration')):
                self.signals[key] = np.c_[self.signals[key], st[k]]
            self.signals['time'] = np.r_[self.signals['time'], t1]


class Obstacle2D(ObstaclexD):

    def __init__(self, initial, shape, simulation, options):
        ObstaclexD.__init__(self, initial, shape, simulation, options)

    def set_default_options(self):
        ObstaclexD.set_default_options(self)
        self.options['horizon_time'] = None

    def init(self, horizon_times=None):
        ObstaclexD.init(self, horizon_times=horizon_times)
        if self.signals['angular_velocity'][:, -1] == 0.:
            self.cos = np.cos(self.signals['orientation'][:, -1][0])
            self.sin = np.sin(self.signals['orientation'][:, -1][0])
            self.gon_weight = 1.
            return
        theta = self.define_parameter('theta', 1)
        omega = self.signals['angular_velocity'][:, -1][0]
        theta0 = theta - self.t * omega
        Ts = 2. * np.pi / abs(omega)
        if self.options['horizon_time'] is None:
            raise ValueError(
                'You need to provide a horizon time when using rotating obstacles!')
        T = self.options['horizon_time']
        n_quarters = int(np.ceil(4 * T / Ts))
        knots_theta = np.r_[np.zeros(3), np.hstack(
            [0.25 * k * np.ones(2) for k in range(1, n_quarters + 1)]),
            0.25 * n_quarters] * (Ts / T)
        Tf, knots = get_interval_T(BSplineBasis(knots_theta, 2), 0, 1.)
        basis = BSplineBasis(knots, 2)
        # NURBS circle: quarter arcs, middle weights sqrt(2)/2
        r = np.sqrt(2.) / 2.
        cos_cfs = np.r_[1., r, 0., -r, -1., -r, 0., r, 1.]
        sin_cfs = np.r_[0., r, 1., r, 0., -r, -1., -r, 0.]
        weight_cfs = np.r_[1., r, 1., r, 1., r, 1., r, 1.]
        n_full = Tf.shape[1]
        cos_cfs = Tf.dot(np.array([cos_cfs[k % 8] for k in range(n_full)]))
        sin_cfs = Tf.dot(np.array([sin_cfs[k % 8] for k in range(n_full)]))
        weight_cfs = Tf.dot(np.array([weight_cfs[k % 8] for k in range(n_full)]))
        cos_wt = BSpline(basis, cos_cfs)
        sin_wt = BSpline(basis, sin_cfs) * np.sign(omega)
        self.cos = cos_wt * cos(theta0) - sin_wt * sin(theta0)
        self.sin = cos_wt * sin(theta0) + sin_wt * cos(theta0)
        self.gon_weight = BSpline(basis, weight_cfs)

    def define_collision_constraints(self, hyperplanes):
        nd = self.n_dim
        for hyperplane in hyperplanes:
            a, b = hyperplane['a'], hyperplane['b']
            for l in range(self.checkpoints.shape[0] // nd):
                xpos = self.pos_spline[0] * self.gon_weight + \
                    self.checkpoints[l * nd + 0] * self.cos - \
                    self.checkpoints[l * nd + 1] * self.sin
                ypos = self.pos_spline[1] * self.gon_weight + \
                    self.checkpoints[l * nd + 0] * self.sin + \
                    self.checkpoints[l * nd + 1] * self.cos
                self.define_constraint(
                    -(a[0] * xpos + a[1] * ypos) +
                    self.gon_weight * (b + self.rad[l]), -inf, 0.)

    def set_parameters(self, current_time):
        parameters = ObstaclexD.set_parameters(self, current_time)
        if 'theta' in self._parameters:
            parameters[self]['theta'] = self.signals['orientation'][:, -1]
        return parameters

    def set_state(self, dictionary):
        ObstaclexD.set_state(self, dictionary)
        for key in ['orientation', 'angular_velocity']:
            if key in dictionary:
                self.signals[key] = np.c_[dictionary[key]]
            else:
                self.signals[key] = np.zeros((1, 1))

    def prepare_simulation(self, initial, simulation):
        ObstaclexD.prepare_simulation(self, initial, simulation)
        for key in ['orientation', 'angular_velocity']:
            if key in initial:
                self.signals[key] = np.c_[initial[key]].astype(float)
            else:
                self.signals[key] = np.zeros((1, 1))

    def simulate(self, simulation_time, sample_time):
        ObstaclexD.simulate(self, simulation_time, sample_time)
        n_samp = int(np.round(simulation_time / sample_time, 6))
        for _ in range(n_samp):
            theta0 = self.signals['orientation'][:, -1][0]
            omega0 = self.signals['angular_velocity'][:, -1][0]
            self.signals['orientation'] = np.c_[
                self.signals['orientation'], theta0 + sample_time * omega0]
            self.signals['angular_velocity'] = np.c_[
                self.signals['angular_velocity'], omega0]


class Obstacle3D(ObstaclexD):

    def __init__(self, initial, shape, simulation, options):
        ObstaclexD.__init__(self, initial, shape, simulation, options)

    def define_collision_constraints(self, hyperplanes):
        nd = self.n_dim
        for hyperplane in hyperplanes:
            a, b = hyperplane['a'], hyperplane['b']
            for l in range(self.checkpoints.shape[0] // nd):
                self.define_constraint(
                    -sum([a[k] * (self.checkpoints[l * nd + k] + self.pos_spline[k])
                          for k in range(nd)]) + b + self.rad[l], -inf, 0.)
