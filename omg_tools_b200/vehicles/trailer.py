"""Trailer pulled by a Dubins vehicle: one more flat output, tan(theta_trailer / 2), next to
the lead vehicle's v~ and tan(theta/2); the trailer's turn rate follows the hitch kinematics
2 d(tg_tr) l = T v~ (2 tg (1 - tg_tr^2) - (1 - tg^2) 2 tg_tr) within a band of 1e-3, the
articulation angle is limited, and the trailer's shape is checked at the hitch offset behind
the lead vehicle's position (reference ``omgtools/vehicles/trailer.py``: constructor 30-40,
trajectory constraints 48-68, initial / terminal constraints 70-96, parameters 146-163,
collision constraints 165-172, signals 174-190, ode 192-206).

As in the reference's example (examples/p2p_trailer.py:31-48) the lead vehicle is ALSO a
child and a vehicle of the problem (``problem.father.add(vehicle)``,
``problem.vehicles.append(vehicle)``): its parameters and rows live under its own label, once
for its own splines and once for the trailer's copies of them."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/trailer.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .dubins import Dubins
from .vehicle import Vehicle
from ..basics.optilayer import inf
from ..basics.shape import Circle
from ..basics.spline_extra import sample_splines


class Trailer(Vehicle):

    def __init__(self, lead_veh=None, shapes=None, l_hitch=0.2, options=None, bounds=None):
        bounds = bounds or {}
        self.lead_veh = Dubins(Circle(0.2)) if lead_veh is None else lead_veh
        Vehicle.__init__(self, n_spl=1 + self.lead_veh.n_spl, degree=3,
                         shapes=shapes if shapes is not None else Circle(0.2), options=options)
        self.l_hitch = l_hitch
        self.tmax = bounds.get('tmax', np.pi / 4.)      # articulation angle limits
        self.tmin = bounds.get('tmin', -np.pi / 4.)

    def init(self):
        self.lead_veh.init()

    def define_trajectory_constraints(self, splines, horizon_time):
        T = horizon_time
        tg_ha_tr = splines[0]
        dtg_ha_tr = tg_ha_tr.derivative()
        v_til_veh, tg_ha_veh = splines[1:]
        eps = 1e-3
        hitch = T * v_til_veh * (2 * tg_ha_veh * (1 - tg_ha_tr**2) - (1 - tg_ha_veh**2) * 2 * tg_ha_tr)
        self.define_constraint(2 * dtg_ha_tr * self.l_hitch - hitch - T * eps, -inf, 0.)
        self.define_constraint(-2 * dtg_ha_tr * self.l_hitch + hitch - T * eps, -inf, 0.)
        self.define_constraint(tg_ha_veh - tg_ha_tr - np.tan(self.tmax / 2.), -inf, 0.)
        self.define_constraint(-tg_ha_veh + tg_ha_tr + np.tan(self.tmin / 2.), -inf, 0.)
        self.lead_veh.define_trajectory_constraints(splines[1:], T)

    def get_initial_constraints(self, splines, horizon_time):
        tg_ha_tr0 = self.define_parameter('tg_ha_tr0', 1)
        dtg_ha_tr0 = self.define_parameter('dtg_ha_tr0', 1)
        tg_ha_tr = splines[0]
        # (the reference returns a 4-tuple here, trailer.py:78: its caller unpacks the first
        # two entries -- the initial condition on d(tg_tr) is never imposed)
        con_tr = [(tg_ha_tr, tg_ha_tr0)]
        return con_tr + self.lead_veh.get_initial_constraints(splines[1:], horizon_time)

    def get_terminal_constraints(self, splines, horizon_time=None):
        if hasattr(self, 'theta_trT'):
            tg_ha_trT = self.define_parameter('tg_ha_trT', 1)
            term_con_tr = [(splines[0], tg_ha_trT)]
        else:
            term_con_tr = []
        con_veh = self.lead_veh.get_terminal_constraints(splines[1:], horizon_time)
        return [term_con_tr + con_veh[0], con_veh[1]]

    def set_initial_conditions(self, state, input=None):
        full = np.zeros(6)
        full[2] = state                                   # trailer heading, set by the user
        full[3:] = self.lead_veh.prediction['state']      # the lead vehicle comes first
        self.prediction['state'] = full
        self.prediction['input'] = self.lead_veh.prediction['input']

    def set_terminal_conditions(self, theta):
        self.theta_trT = theta

    def get_init_spline_value(self, subgoals=None):
        L = len(self.basis)
        tg0 = np.tan(self.prediction['state'][2] / 2.)
        tgT = np.tan(self.theta_trT / 2.) if hasattr(self, 'theta_trT') else tg0
        return [np.c_[np.linspace(tg0, tgT, L), self.lead_veh.get_init_spline_value()[0]]]

    def check_terminal_conditions(self):
        tol = self.options['stop_tol']
        ok = True
        if hasattr(self, 'theta_trT'):
            ok = np.linalg.norm(self.signals['state'][2, -1] - self.theta_trT) <= tol
        return self.lead_veh.check_terminal_conditions() and ok

    def set_parameters(self, current_time):
        self.lead_veh.prediction['input'] = self.prediction['input']
        self.lead_veh.prediction['state'] = self.prediction['state'][3:]
        parameters = Vehicle.set_parameters(self, current_time)
        st, inp = self.prediction['state'], self.prediction['input']
        p = {'tg_ha_tr0': np.tan(st[2] / 2.)}
        p['dtg_ha_tr0'] = 0.5 * inp[0] / self.l_hitch * np.sin(st[5] - st[2]) * (1 + p['tg_ha_tr0']**2)
        if hasattr(self, 'theta_trT'):
            p['tg_ha_trT'] = np.tan(self.theta_trT / 2.)
        parameters[self].update(p)
        parameters[self].update(self.lead_veh.set_parameters(current_time)[self.lead_veh])
        return parameters

    def define_collision_constraints(self, hyperplanes, room, splines, horizon_time):
        tg_ha_tr = splines[0]
        x_veh, y_veh = self.lead_veh._flat_position(splines[1:], horizon_time) \
            if not self.lead_veh.options['substitution'] else (self.lead_veh.x, self.lead_veh.y)
        # the trailer sits l_hitch behind the lead vehicle's reference point
        self.define_collision_constraints_2d(hyperplanes, room, [x_veh, y_veh], horizon_time,
                                             tg_ha=tg_ha_tr, offset=-self.l_hitch)
        self.lead_veh.define_collision_constraints(hyperplanes, room, splines[1:], horizon_time)

    def splines2signals(self, splines, time):
        signals = {}
        tg_ha_tr = splines[0]
        dtg_ha_tr = tg_ha_tr.derivative()
        tg = np.array(sample_splines([tg_ha_tr], time))
        dtg = np.array(sample_splines([dtg_ha_tr], time))
        theta_tr = 2 * np.arctan2(tg, 1)
        self.lead_veh.prediction['state'] = self.prediction['state'][3:]
        sv = self.lead_veh.splines2signals(splines[1:], time)
        x_tr = sv['state'][0, :] - self.l_hitch * np.cos(theta_tr)
        y_tr = sv['state'][1, :] - self.l_hitch * np.sin(theta_tr)
        signals['state'] = np.r_[x_tr, y_tr, theta_tr, sv['state']]
        signals['pose'] = signals['state']
        signals['input'] = sv['input']
        signals['r1'] = np.r_[tg, dtg]
        return signals

    def state2pose(self, state):
        return np.r_[state[:3], self.lead_veh.state2pose(state[3:])]

    def ode(self, state, input):
        _, _, theta_tr, x_veh, y_veh, theta_veh = state
        dtheta_tr = input[0] / self.l_hitch * np.sin(theta_veh - theta_tr)
        ode_veh = self.lead_veh.ode([x_veh, y_veh, theta_veh], input)
        return np.r_[ode_veh[0] + self.l_hitch * np.sin(theta_tr) * dtheta_tr,
                     ode_veh[1] - self.l_hitch * np.cos(theta_tr) * dtheta_tr,
                     dtheta_tr, ode_veh]
