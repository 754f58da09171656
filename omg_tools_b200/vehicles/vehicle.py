"""Vehicle base class: spline definition and collision/room constraint rows.

Model-building interface and row order follow the reference's
``omgtools/vehicles/vehicle.py`` (define_knots 80-87, define_splines 105-120,
define_collision_constraints_2d 122-190, _3d 192-232, get_fleet_center 234-241).
Host-side bookkeeping after a solve (store / predict / simulate) is reduced to
the ideal, noise-free case: the vehicle follows its computed spline exactly
(reference options 'ideal_prediction'/'ideal_update', vehicle.py:70-78,
302-337, 339-410); plotting, delays and disturbances are out of scope.
"""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/vehicle.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from ..basics.optilayer import OptiChild, inf
from ..basics.poly import rel_time
from ..basics.spline import BSplineBasis
from ..basics.spline_extra import definite_integral, sample_splines
from ..basics.shape import Rectangle, Square, Circle


class Vehicle(OptiChild):

    def __init__(self, n_spl, degree, shapes, options=None):
        options = options or {}
        OptiChild.__init__(self, 'vehicle')
        self.shapes = shapes if isinstance(shapes, list) else [shapes]
        self.n_dim = self.shapes[0].n_dim
        for shape in self.shapes:
            if shape.n_dim != self.n_dim:
                raise ValueError('All vehicle shapes should have same spatial' +
                                 'dimension.')
        self.prediction = {}
        self.init_spline_values = None
        self.degree = degree
        self.to_simulate = True
        self.set_default_options()
        self.set_options(options)
        self.define_knots(knot_intervals=10)
        self.n_spl = n_spl

    # ------------------------------------------------------------------
    # options
    # ------------------------------------------------------------------

    def set_default_options(self):
        self.options = {'safety_distance': 0., 'safety_weight': 10.,
                        'room_constraints': True, 'stop_tol': 1.e-3,
                        'ideal_prediction': True, 'ideal_update': True,
                        '1storder_delay': False, 'time_constant': 0.1,
                        'input_disturbance': None}

    def set_options(self, options):
        self.options.update(options)

    def define_knots(self, **kwargs):
        if 'knot_intervals' in kwargs:
            self.knot_intervals = kwargs['knot_intervals']
            self.knots = np.r_[np.zeros(self.degree), np.linspace(
                0., 1., self.knot_intervals + 1), np.ones(self.degree)]
        if 'knots' in kwargs:
            self.knots = kwargs['knots']
        self.basis = BSplineBasis(self.knots, self.degree)

    def set_init_spline_values(self, values, n_seg=1):
        self.init_spline_values = [0] * n_seg
        for k in range(n_seg):
            if values[k].shape != (len(self.basis), self.n_spl):
                raise ValueError('Initial guess has wrong dimensions for spline ' +
                                 str(k) + ', required: ' +
                                 str((len(self.basis), self.n_spl)) +
                                 ' while you gave: ' + str(values[k].shape))
            self.init_spline_values[k] = values[k]

    # ------------------------------------------------------------------
    # optimization modelling
    # ------------------------------------------------------------------

    def define_splines(self, n_seg=1):
        self.n_seg = n_seg
        self.splines = []
        if self.init_spline_values is not None:
            init = self.init_spline_values
            self.init_spline_values = None
        else:
            init = [None] * n_seg
        for k in range(self.n_seg):
            spline = self.define_spline_variable(
                'splines_seg' + str(k), self.n_spl, value=init[k])
            self.splines.append(spline)
        return self.splines

    def _share_spline(self, name, spline):
        """Spline whose non-linear coefficients are named intermediates (basics/poly.py
        new_mid): products that many rows repeat -- (1 +- tg^2), position x (1 + tg^2) of a
        vehicle with a heading -- are then evaluated once and the rows are bilinear in the
        hyperplane coefficients and the intermediates, instead of expanding every product
        spline into every row (the shared nodes of CasADi's expression graph).  Coefficients
        that already contain intermediates are left as they are (no nesting)."""
        from ..basics.poly import Poly, new_mid, sym_info, resolve
        from ..basics.spline import BSpline
        if not isinstance(spline, BSpline):
            return spline
        cnt = self.__dict__.setdefault('_share_cnt', [0])
        coeffs = np.empty(len(spline.coeffs), dtype=object)
        for k, c in enumerate(spline.coeffs):
            if (isinstance(c, Poly) and c.degree() >= 2 and
                    not any(sym_info(resolve(sid)).kind == 'mid' for sid in c.symbols())):
                c = new_mid('%s_%s%d_%d' % (self.label, name, cnt[0], k), c)
            coeffs[k] = c
        cnt[0] += 1
        return BSpline(spline.basis, coeffs)

    def define_collision_constraints_2d(self, hyperplanes, room, positions,
                                        horizon_time, tg_ha=0, offset=0):
        t = self.define_symbol('t')
        safety_distance = self.options['safety_distance']
        safety_weight = self.options['safety_weight']
        positions = [positions] if not isinstance(positions[0], list) \
            else positions
        from ..basics.spline import BSpline
        heading = isinstance(tg_ha, BSpline)
        if heading:     # products shared by every row of this call (see _share_spline)
            one_m, one_p = self._share_spline('c', 1. - tg_ha**2), self._share_spline('q', 1 + tg_ha**2)
        for s, shape in enumerate(self.shapes):
            position = positions[s]
            checkpoints, rad = shape.get_checkpoints()
            if heading:
                posq = [self._share_spline('px', position[0] * (1 + tg_ha**2) + offset * (1 - tg_ha**2)),
                        self._share_spline('py', position[1] * (1 + tg_ha**2) + offset * (2 * tg_ha))]
            # obstacle avoidance: a.(R chk + pos) - b + rad + sd - eps <= 0
            if shape in hyperplanes:
                for k, hyperplane in enumerate(hyperplanes[shape]):
                    a, b = hyperplane['a'], hyperplane['b']
                    sl = 1 if 'slack' not in hyperplane else hyperplane['slack']
                    if safety_distance > 0.:
                        eps = self.define_spline_variable(
                            'eps_' + str(s) + str(k))[0]
                        obj = safety_weight * definite_integral(
                            eps, rel_time(t, horizon_time), 1.)
                        self.define_objective(obj)
                        self.define_constraint(eps - safety_distance, -inf, 0.)
                        self.define_constraint(-eps, -inf, 0.)
                    else:
                        eps = 0.
                    for l, chck in enumerate(checkpoints):
                        con = 0
                        if heading:
                            con += (a[0] * chck[0] + a[1] * chck[1]) * one_m
                            con += (-a[0] * chck[1] + a[1] * chck[0]) * (2 * tg_ha)
                            con += (a[0] * posq[0] + a[1] * posq[1])
                            con += (-b + sl * rad[l] + safety_distance - eps) * one_p
                        else:
                            con += (a[0] * chck[0] + a[1] * chck[1]) * (1. - tg_ha**2)
                            con += (-a[0] * chck[1] + a[1] * chck[0]) * (2 * tg_ha)
                            pos = [0, 0]
                            pos[0] = position[0] * (1 + tg_ha**2) + offset * (1 - tg_ha**2)
                            pos[1] = position[1] * (1 + tg_ha**2) + offset * (2 * tg_ha)
                            con += (a[0] * pos[0] + a[1] * pos[1])
                            con += (-b + sl * rad[l] + safety_distance - eps) * (1 + tg_ha**2)
                        self.define_constraint(con, -inf, 0)
            # room constraints
            if self.options['room_constraints']:
                lims = room['shape'].get_canvas_limits()
                room_limits = [lims[k] + room['position'][k]
                               for k in range(self.n_dim)]
                axis_aligned = (
                    isinstance(room['shape'], (Rectangle, Square)) and
                    room['shape'].orientation == 0.0 and
                    (isinstance(shape, Circle) or
                     (isinstance(shape, (Rectangle, Square)) and
                      shape.orientation == 0)) and
                    isinstance(tg_ha, (int, float)) and tg_ha == 0.)
                if axis_aligned:
                    for chck in checkpoints:
                        for k in range(self.n_dim):
                            self.define_constraint(
                                -(chck[k] + position[k]) + room_limits[k][0] + rad[0], -inf, 0.)
                            self.define_constraint(
                                (chck[k] + position[k]) - room_limits[k][1] + rad[0], -inf, 0.)
                else:
                    hyp_room = room['shape'].get_hyperplanes(
                        position=room['position'])
                    for l, chck in enumerate(checkpoints):
                        for hpp in hyp_room.values():
                            con = 0
                            if heading:
                                con += (hpp['a'][0] * chck[0] + hpp['a'][1] * chck[1]) * one_m
                                con += (-hpp['a'][0] * chck[1] + hpp['a'][1] * chck[0]) * (2 * tg_ha)
                                con += (hpp['a'][0] * posq[0] + hpp['a'][1] * posq[1])
                                con += (-hpp['b'] + rad[l]) * one_p
                            else:
                                con += (hpp['a'][0] * chck[0] + hpp['a'][1] * chck[1]) * (1. - tg_ha**2)
                                con += (-hpp['a'][0] * chck[1] + hpp['a'][1] * chck[0]) * (2 * tg_ha)
                                pos = [0, 0]
                                pos[0] = position[0] * (1 + tg_ha**2) + offset * (1 - tg_ha**2)
                                pos[1] = position[1] * (1 + tg_ha**2) + offset * (2 * tg_ha)
                                con += (hpp['a'][0] * pos[0] + hpp['a'][1] * pos[1])
                                con += (-hpp['b'] + rad[l]) * (1 + tg_ha**2)
                            self.define_constraint(con, -inf, 0)

    def define_collision_constraints_3d(self, hyperplanes, room, positions,
                                        horizon_time):
        t = self.define_symbol('t')
        safety_distance = self.options['safety_distance']
        safety_weight = self.options['safety_weight']
        positions = [positions] if not isinstance(positions[0], list) \
            else positions
        for s, shape in enumerate(self.shapes):
            position = positions[s]
            checkpoints, rad = shape.get_checkpoints()
            if shape in hyperplanes:
                for k, hyperplane in enumerate(hyperplanes[shape]):
                    a, b = hyperplane['a'], hyperplane['b']
                    if safety_distance > 0.:
                        eps = self.define_spline_variable(
                            'eps_' + str(s) + str(k))[0]
                        obj = safety_weight * definite_integral(
                            eps, rel_time(t, horizon_time), 1.)
                        self.define_objective(obj)
                        self.define_constraint(eps - safety_distance, -inf, 0.)
                        self.define_constraint(-eps, -inf, 0.)
                    else:
                        eps = 0.
                    for l, chck in enumerate(checkpoints):
                        self.define_constraint(
                            sum([a[q] * (chck[q] + position[q]) for q in range(3)])
                            - b + rad[l] + safety_distance - eps, -inf, 0)
            if self.options['room_constraints']:
                lims = room['shape'].get_canvas_limits()
                room_limits = [lims[k] + room['position'][k]
                               for k in range(self.n_dim)]
                for chck in checkpoints:
                    for k in range(3):
                        self.define_constraint(
                            -(chck[k] + position[k]) + room_limits[k][0], -inf, 0.)
                        self.define_constraint(
                            (chck[k] + position[k]) - room_limits[k][1], -inf, 0.)

    def get_fleet_center(self, splines, rel_pos, substitute=True):
        rel_pos = list(rel_pos) if not isinstance(rel_pos, list) else rel_pos
        center = [s + rp for s, rp in zip(splines, rel_pos)]
        if substitute:
            return self.define_substitute('fleet_center', center)
        return center

    def set_parameters(self, current_time):
        return {self: {}}

    def init(self):
        pass

    # ------------------------------------------------------------------
    # ideal host-side bookkeeping between MPC steps
    # ------------------------------------------------------------------

    def store(self, current_time, sample_time, spline_segments, segment_times,
              time_axis=None, **kwargs):
        """Sample the computed splines into trajectories
        (reference vehicle.py:250-300, single segment)."""
        splines = spline_segments[0]
        horizon_time = segment_times if not isinstance(segment_times, list) \
            else segment_times[0]
        if time_axis is None:
            n_samp = int(round(horizon_time / sample_time, 6)) + 1
            time_axis = np.linspace(0., (n_samp - 1) * sample_time, n_samp)
        self.trajectories = self.splines2signals(
            [s.scale(horizon_time) for s in splines], time_axis)
        self.trajectories['time'] = time_axis - time_axis[0] + current_time
        self.trajectories['splines'] = np.c_[
            sample_splines([s.scale(horizon_time) for s in splines], time_axis)]
        if not hasattr(self, 'signals'):
            self.signals = {}
            for key in self.trajectories:
                val = np.atleast_2d(self.trajectories[key])
                self.signals[key] = val[:, :1] if key != 'time' \
                    else np.array([[current_time]])

    def predict(self, current_time, predict_time, sample_time, state0=None,
                input0=None, dinput0=None, delay=0, enforce_states=False,
                enforce_inputs=False):
        """Where the vehicle will be when the next trajectory starts (reference vehicle.py:302-337).

        enforce_states (+ enforce_inputs): the measured state0 (and input0) -- or, without a
        measurement, the last simulated signal -- becomes the initial condition as is.
        Otherwise the stored trajectory is read predict_time (+ ``delay`` samples of computation
        delay) ahead.  Only the ideal prediction is done on the host; the non-ideal one (ODE
        integration of the planned inputs) is the batched RK4 kernel of execution/batch_mpc.py."""
        last = lambda key: np.atleast_2d(self.signals[key])[:, -1]
        if enforce_states and enforce_inputs:
            if state0 is not None and input0 is not None:
                self.set_initial_conditions(state0, input0)
            elif hasattr(self, 'signals') and 'state' in self.signals:
                self.set_initial_conditions(last('state'), last('input'))
            return
        if enforce_states:
            if state0 is not None:
                self.set_initial_conditions(state0)
            elif hasattr(self, 'signals') and 'state' in self.signals:
                self.set_initial_conditions(last('state'))
            return
        if not hasattr(self, 'trajectories'):
            return
        if not self.options.get('ideal_prediction', True):
            raise NotImplementedError('non-ideal prediction on the host: use execution.batch_mpc '
                                      '(omg_integrate_rk4) or ideal_prediction=True')
        n_samp = int(np.round(predict_time / sample_time, 6)) + int(delay)
        for key, val in self.trajectories.items():
            # every sampled signal is predicted (state, input, and model specific
            # ones such as the quadrotor's dspl / ddspl)
            if key not in ('time', 'splines'):
                self.prediction[key] = np.atleast_2d(val)[:, n_samp]

    def simulate(self, simulation_time, sample_time):
        """Ideal update: follow the stored trajectory."""
        n_samp = int(np.round(simulation_time / sample_time, 6))
        for key in self.trajectories:
            val = np.atleast_2d(self.trajectories[key])
            self.signals[key] = np.c_[self.signals[key], val[:, 1:n_samp + 1]]
