"""Holonomic vehicle on a line: one degree-3 spline, velocity/acceleration
bounds, no collision rows (reference ``omgtools/vehicles/holonomic1d.py``:
bounds 31-36, trajectory constraints 41-49, initial/terminal constraints
51-67, initial guess 78-83, parameters 94-99)."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/holonomic1d.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .vehicle import Vehicle
from ..basics.optilayer import inf
from ..basics.shape import Rectangle
from ..basics.spline_extra import sample_splines


class Holonomic1D(Vehicle):

    def __init__(self, width=0.7, height=0.1, options=None, bounds=None):
        bounds = bounds or {}
        Vehicle.__init__(self, n_spl=1, degree=3,
                         shapes=Rectangle(width, height), options=options)
        self.vmin = bounds.get('vmin', -0.5)
        self.vmax = bounds.get('vmax', 0.5)
        self.amin = bounds.get('amin', -1.)
        self.amax = bounds.get('amax', 1.)

    def define_trajectory_constraints(self, splines, horizon_time):
        T = horizon_time
        x = splines[0]
        dx, ddx = x.derivative(), x.derivative(2)
        self.define_constraint(-dx + T * self.vmin, -inf, 0.)
        self.define_constraint(dx - T * self.vmax, -inf, 0.)
        self.define_constraint(-ddx + (T**2) * self.amin, -inf, 0.)
        self.define_constraint(ddx - (T**2) * self.amax, -inf, 0.)

    def get_initial_constraints(self, splines, horizon_time):
        state0 = self.define_parameter('state0')
        input0 = self.define_parameter('input0')
        x = splines[0]
        # 1x1 parameters are scalars in this layer (CasADi needs the [0])
        return [(x, state0), (x.derivative(), horizon_time * input0)]

    def get_terminal_constraints(self, splines, horizon_time=None):
        position = self.define_parameter('poseT')
        x = splines[0]
        term_con = [(x, position)]
        term_con_der = [(x.derivative(d), 0.) for d in range(1, self.degree + 1)]
        return [term_con, term_con_der]

    def set_initial_conditions(self, state, input=None):
        self.prediction['state'] = np.atleast_1d(np.asarray(state, dtype=float))
        self.prediction['input'] = np.atleast_1d(
            np.asarray(0. if input is None else input, dtype=float))

    def set_terminal_conditions(self, position):
        self.poseT = np.atleast_1d(np.asarray(position, dtype=float))

    def get_init_spline_value(self, subgoals=None):
        pos0, posT = self.prediction['state'][0], self.poseT[0]
        return [np.linspace(pos0, posT, len(self.basis)).reshape(-1, 1)]

    def check_terminal_conditions(self):
        tol = self.options['stop_tol']
        if (np.linalg.norm(self.signals['state'][:, -1] - self.poseT) > tol or
                np.linalg.norm(self.signals['input'][:, -1]) > tol):
            return False
        return True

    def set_parameters(self, current_time):
        parameters = Vehicle.set_parameters(self, current_time)
        parameters[self]['state0'] = self.prediction['state']
        parameters[self]['input0'] = self.prediction['input']
        parameters[self]['poseT'] = self.poseT
        return parameters

    def define_collision_constraints(self, hyperplanes, room, splines, horizon_time):
        pass

    def splines2signals(self, splines, time):
        signals = {}
        x = splines[0]
        dx, ddx = x.derivative(), x.derivative(2)
        signals['state'] = np.c_[sample_splines([x], time)]
        signals['input'] = np.c_[sample_splines([dx], time)]
        signals['a'] = np.c_[sample_splines([ddx], time)]
        return signals

    def state2pose(self, state):
        return np.r_[state, 0., 0.]

    def ode(self, state, input):
        return input
