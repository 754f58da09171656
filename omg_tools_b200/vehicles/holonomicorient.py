"""Holonomic vehicle with a free heading: x, y and tan(theta/2) are degree-3 splines; the
heading enters the collision rows of a non-circular shape (vehicle.py:
define_collision_constraints_2d with tg_ha) and the turn-rate limits
2 d(tg)/dt <= (1 + tg^2) w_max (reference ``omgtools/vehicles/holonomicorient.py``: bounds
28-37, trajectory constraints 48-85 incl. the optional regularisation of the heading
rate, initial / terminal constraints 87-108, initial guess 120-135, parameters 145-156,
collision constraints 158-160, signals 162-177)."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/holonomicorient.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .vehicle import Vehicle
from ..basics.optilayer import inf
from ..basics.poly import rel_time
from ..basics.shape import Rectangle
from ..basics.spline_extra import sample_splines, definite_integral


class HolonomicOrient(Vehicle):

    def __init__(self, shapes=None, options=None, bounds=None):
        bounds = bounds or {}
        shapes = shapes if shapes is not None else Rectangle(width=0.2, height=0.4)
        Vehicle.__init__(self, n_spl=3, degree=3, shapes=shapes, options=options)
        self.vmin = bounds.get('vmin', -0.5)
        self.vmax = bounds.get('vmax', 0.5)
        self.amin = bounds.get('amin', -1.)
        self.amax = bounds.get('amax', 1.)
        self.wmin = bounds.get('wmin', -np.pi / 6.)
        self.wmax = bounds.get('wmax', np.pi / 6.)

    def set_default_options(self):
        Vehicle.set_default_options(self)
        self.options.update({'syslimit': 'norm_inf', 'reg_type': None})

    def init(self):
        self.t = self.define_symbol('t')

    def define_trajectory_constraints(self, splines, horizon_time):
        T = horizon_time
        x, y, tg_ha = splines
        dx, dy, dtg_ha = x.derivative(), y.derivative(), tg_ha.derivative()
        ddx, ddy = x.derivative(2), y.derivative(2)
        if self.options['syslimit'] == 'norm_2':
            self.define_constraint((dx**2 + dy**2) - (T**2) * self.vmax**2, -inf, 0.)
            self.define_constraint((ddx**2 + ddy**2) - (T**4) * self.amax**2, -inf, 0.)
        elif self.options['syslimit'] == 'norm_inf':
            self.define_constraint(-dx + T * self.vmin, -inf, 0.)
            self.define_constraint(-dy + T * self.vmin, -inf, 0.)
            self.define_constraint(dx - T * self.vmax, -inf, 0.)
            self.define_constraint(dy - T * self.vmax, -inf, 0.)
            self.define_constraint(-ddx + (T**2) * self.amin, -inf, 0.)
            self.define_constraint(-ddy + (T**2) * self.amin, -inf, 0.)
            self.define_constraint(ddx - (T**2) * self.amax, -inf, 0.)
            self.define_constraint(ddy - (T**2) * self.amax, -inf, 0.)
        else:
            raise ValueError('Only norm_2 and norm_inf are defined as system limit.')
        # turn rate: theta = 2 atan(tg)  =>  dtheta = 2 dtg / (1 + tg^2)
        self.define_constraint(2 * dtg_ha - (1 + tg_ha**2) * T * self.wmax, -inf, 0.)
        self.define_constraint(-2 * dtg_ha + (1 + tg_ha**2) * T * self.wmin, -inf, 0.)
        reg, weight = self.options['reg_type'], self.options.get('reg_weight', 0.0)
        if reg == 'norm_1' and weight != 0.0:
            g_reg = self.define_spline_variable('g_reg', 1, basis=dtg_ha.basis)[0]
            objective = definite_integral(g_reg, rel_time(self.t, T), 1.)
            self.define_constraint(dtg_ha - g_reg, -inf, 0.)
            self.define_constraint(-dtg_ha - g_reg, -inf, 0.)
            self.define_objective(weight * objective)
        if reg == 'norm_2' and weight != 0.0:
            self.define_objective(weight * definite_integral(dtg_ha**2, rel_time(self.t, T), 1.))

    def get_initial_constraints(self, splines, horizon_time):
        pos0 = self.define_parameter('pos0', 2)
        tg_ha0 = self.define_parameter('tg_ha0', 1)
        vel0 = self.define_parameter('vel0', 2)
        dtg_ha0 = self.define_parameter('dtg_ha0', 1)
        x, y, tg_ha = splines
        dx, dy, dtg_ha = x.derivative(), y.derivative(), tg_ha.derivative()
        return [(x, pos0[0]), (y, pos0[1]), (tg_ha, tg_ha0),
                (dx, horizon_time * vel0[0]), (dy, horizon_time * vel0[1]),
                (dtg_ha, horizon_time * dtg_ha0)]

    def get_terminal_constraints(self, splines, horizon_time=None):
        posT = self.define_parameter('posT', 2)
        tg_haT = self.define_parameter('tg_haT', 1)
        x, y, tg_ha = splines
        term_con = [(x, posT[0]), (y, posT[1]), (tg_ha, tg_haT)]
        term_con_der = []
        for d in range(1, self.degree + 1):
            term_con_der.extend([(x.derivative(d), 0.), (y.derivative(d), 0.),
                                 (tg_ha.derivative(d), 0.)])
        return [term_con, term_con_der]

    def set_initial_conditions(self, state, input=None):
        if input is None:
            input = np.zeros(3)
        self.prediction['state'] = np.asarray(state, dtype=float)
        self.prediction['input'] = np.asarray(input, dtype=float)

    def set_terminal_conditions(self, pose):
        self.poseT = np.asarray(pose, dtype=float)

    def get_init_spline_value(self, subgoals=None):
        # the reference interpolates x and y only; the heading spline starts at zero
        # (holonomicorient.py:131-134)
        L = len(self.basis)
        init_value = np.zeros((L, 3))
        for k in range(2):
            init_value[:, k] = np.linspace(self.prediction['state'][k], self.poseT[k], L)
        return [init_value]

    def check_terminal_conditions(self):
        tol = self.options['stop_tol']
        if (np.linalg.norm(self.signals['state'][:, -1] - self.poseT) > tol or
                np.linalg.norm(self.signals['input'][:, -1]) > tol):
            return False
        return True

    def set_parameters(self, current_time):
        parameters = Vehicle.set_parameters(self, current_time)
        p = parameters[self]
        p['pos0'] = self.prediction['state'][:2]
        p['tg_ha0'] = np.tan(self.prediction['state'][2] / 2)
        p['vel0'] = self.prediction['input'][:2]
        p['dtg_ha0'] = 0.5 * self.prediction['input'][2] * (1 + p['tg_ha0']**2)
        p['posT'] = self.poseT[:2]
        p['tg_haT'] = np.tan(self.poseT[2] / 2)
        return parameters

    def define_collision_constraints(self, hyperplanes, room, splines, horizon_time):
        x, y, tg_ha = splines[0], splines[1], splines[2]
        self.define_collision_constraints_2d(hyperplanes, room, [x, y], horizon_time,
                                             tg_ha=tg_ha)

    def splines2signals(self, splines, time):
        signals = {}
        x, y, tg_ha = splines[0], splines[1], splines[2]
        dx, dy, dtg_ha = x.derivative(), y.derivative(), tg_ha.derivative()
        ddx, ddy = x.derivative(2), y.derivative(2)
        tg = np.asarray(sample_splines([tg_ha], time))
        theta = 2 * np.arctan2(tg, 1)
        dtheta = 2 * np.asarray(sample_splines([dtg_ha], time)) / (1 + tg**2)
        inp = np.r_[np.c_[sample_splines([dx, dy], time)], dtheta]
        signals['state'] = np.r_[np.c_[sample_splines([x, y], time)], theta]
        signals['input'] = inp
        signals['v_tot'] = np.sqrt(inp[0, :]**2 + inp[1, :]**2)
        signals['a'] = np.c_[sample_splines([ddx, ddy], time)]
        return signals

    def state2pose(self, state):
        return state

    def ode(self, state, input):
        return input
