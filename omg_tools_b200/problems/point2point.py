"""Point-to-point motion problem with fixed horizon time.

Mirrors the reference's ``omgtools/problems/point2point.py``: construct 53-62,
initial constraints 64-70, terminal constraints + L1 slack objective 151-172,
parameters t/T 174-181, knot-crossing warm-start shift 187-201, store 213-229.
``FreeTPoint2point`` (T as a decision variable, reference 269-374) is
supported for models whose rows stay polynomial in T: no safety-distance
slack (its objective integrates from t/T) and no rotating obstacles.
"""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/problems/point2point.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
from __future__ import print_function

import numpy as np

from .problem import Problem
from ..basics.optilayer import inf
from ..basics.spline_extra import (definite_integral, shiftoverknot_T, evalspline,
                                   shift_spline)


class Point2point(object):
    """Selects between fixed-T and free-T problem (reference 28-35)."""

    def __new__(cls, fleet, environment, options=None, freeT=False):
        if freeT:
            return FreeTPoint2point(fleet, environment, options)
        return FixedTPoint2point(fleet, environment, options)


class Point2pointProblem(Problem):

    def __init__(self, fleet, environment, options):
        Problem.__init__(self, fleet, environment, options, label='p2p')
        self.init_time = None
        self.start_time = 0.

    def set_default_options(self):
        Problem.set_default_options(self)
        self.options['inter_vehicle_avoidance'] = False

    def define_time_symbols(self):
        """T and t are parameters of the fixed-T problem (reference 176-181)."""
        self.T, self.t = self.define_parameter('T'), self.define_parameter('t')
        self.t0 = self.t / self.T

    def construct(self):
        self.define_time_symbols()
        Problem.construct(self)
        for vehicle in self.vehicles:
            splines = vehicle.define_splines(n_seg=1)
            vehicle.define_trajectory_constraints(splines[0], self.T)
            self.environment.define_collision_constraints(vehicle, splines, self.T)
        if len(self.vehicles) > 1 and self.options.get('inter_vehicle_avoidance'):
            self.environment.define_intervehicle_collision_constraints(self.vehicles, self.T)

    def define_init_constraints(self):
        for vehicle in self.vehicles:
            init_con = vehicle.get_initial_constraints(vehicle.splines[0], self.T)
            for spline, condition in init_con:
                self.define_constraint(
                    evalspline(spline, self.t0) - condition, 0., 0.)

    def initialize(self, current_time):
        self.start_time = current_time

    def reinitialize(self, father=None):
        if father is None:
            father = self.father
        Problem.reinitialize(self)
        for vehicle in self.vehicles:
            init = vehicle.get_init_spline_value()
            for k in range(vehicle.n_seg):
                father.set_variables(init[k], vehicle, 'splines_seg' + str(k))

    def set_init_time(self, time):
        self.init_time = time

    def reset_init_time(self):
        self.init_time = None

    def stop_criterium(self, current_time, update_time):
        stop = True
        for vehicle in self.vehicles:
            stop *= vehicle.check_terminal_conditions()
        return stop

    def final(self):
        self.reset_init_time()
        obj = self.compute_objective()
        if self.options['verbose'] >= 1:
            print('\nWe reached our target!')
            print('%-18s %6g' % ('Objective:', obj))
            print('%-18s %6g ms' % ('Max update time:',
                                    max(self.update_times) * 1000.))
            print('%-18s %6g ms' % ('Av update time:',
                                    (sum(self.update_times) * 1000. /
                                     len(self.update_times))))


class FixedTPoint2point(Point2pointProblem):

    def __init__(self, fleet, environment, options):
        Point2pointProblem.__init__(self, fleet, environment, options)
        self.objective = 0.
        if self.vehicles[0].knot_intervals is None:
            raise ValueError('A constant knot interval should be used for ' +
                             'a fixed T point2point problem.')
        self.knot_time = (int(self.options['horizon_time'] * 1000.) /
                          self.vehicles[0].knot_intervals) / 1000.

    def set_default_options(self):
        Point2pointProblem.set_default_options(self)
        self.options['horizon_time'] = 10.
        self.options['hard_term_con'] = False
        self.options['no_term_con_der'] = False

    def construct(self):
        Point2pointProblem.construct(self)
        self.define_init_constraints()
        self.define_terminal_constraints()

    def define_terminal_constraints(self):
        objective = 0.
        self.term_con_len = []
        for vehicle in self.vehicles:
            term_con, term_con_der = vehicle.get_terminal_constraints(
                vehicle.splines[0])
            if self.options.get('no_term_con_der'):
                term_con_der = []
            self.term_con_len.append(len(term_con))
            for k, (spline, condition) in enumerate(term_con):
                g = self.define_spline_variable(
                    'g' + str(k), 1, basis=spline.basis)[0]
                objective += definite_integral(g, self.t0, 1.)
                self.define_constraint(spline - condition - g, -inf, 0.)
                self.define_constraint(-spline + condition - g, -inf, 0.)
                if self.options['hard_term_con']:
                    self.define_constraint(spline(1.) - condition, 0., 0.)
            for spline, condition in term_con_der:
                self.define_constraint(spline(1.) - condition, 0., 0.)
        self.define_objective(objective)

    def set_parameters(self, current_time):
        parameters = Point2pointProblem.set_parameters(self, current_time)
        if self.init_time is None:
            parameters[self]['t'] = np.round(current_time, 6) % self.knot_time
        else:
            parameters[self]['t'] = self.init_time
        parameters[self]['T'] = self.options['horizon_time']
        return parameters

    def init_step(self, current_time, update_time):
        if not hasattr(self, 'current_time_prev'):
            self.current_time_prev = 0
        interval_prev = int(np.round(self.current_time_prev / self.knot_time, 6))
        interval_now = int(np.round(current_time / self.knot_time, 6))
        if interval_prev < interval_now:     # passed a knot
            self.father.transform_primal_splines(
                lambda coeffs, basis, T: T.dot(coeffs))
        self.current_time_prev = current_time

    def init_primal_transform(self, basis):
        return shiftoverknot_T(basis)

    def initialize(self, current_time):
        Point2pointProblem.initialize(self, current_time)
        self.current_time_prev = current_time

    def store(self, current_time, update_time, sample_time):
        horizon_time = self.options['horizon_time']
        if self.init_time is None:
            rel_current_time = np.round(
                current_time - self.start_time, 6) % self.knot_time
        else:
            rel_current_time = self.init_time
        for vehicle in self.vehicles:
            n_samp = int(round((horizon_time - rel_current_time) / sample_time, 6)) + 1
            time_axis = np.linspace(
                rel_current_time, rel_current_time + (n_samp - 1) * sample_time, n_samp)
            spline_segments = [self.father.get_variables(
                vehicle, 'splines_seg' + str(k)) for k in range(vehicle.n_seg)]
            vehicle.store(current_time, sample_time, spline_segments,
                          horizon_time, time_axis)

    def simulate(self, current_time, simulation_time, sample_time):
        horizon_time = self.options['horizon_time']
        if self.init_time is None:
            rel_current_time = np.round(
                current_time - self.start_time, 6) % self.knot_time
        else:
            rel_current_time = self.init_time
        if horizon_time - rel_current_time < simulation_time:
            simulation_time = horizon_time - rel_current_time
        Problem.simulate(self, current_time, simulation_time, sample_time)

    def compute_objective(self):
        obj = 0.
        for v, vehicle in enumerate(self.vehicles):
            for k in range(self.term_con_len[v]):
                g = self.father.get_variables(self, 'g' + str(k))[0]
                obj += self.options['horizon_time'] * g.integral()
        return obj


class FreeEndPoint2point(FixedTPoint2point):
    """Fixed-T point-to-point problem whose terminal conditions (all, or those listed in
    ``free_ind[vehicle]``) are decision variables ``conT<l>`` instead of parameters: the
    agents of a RendezVous agree on them (reference point2point.py:377-417)."""

    def __init__(self, fleet, environment, options, free_ind=None):
        FixedTPoint2point.__init__(self, fleet, environment, options)
        self.free_ind = free_ind

    def construct(self):
        if self.free_ind is None:
            self._all_free = True
            self.free_ind = {}
        FixedTPoint2point.construct(self)

    def define_terminal_constraints(self):
        objective = 0.
        self.term_con_len = []
        spline = condition = None
        for l, vehicle in enumerate(self.vehicles):
            term_con, term_con_der = vehicle.get_terminal_constraints(vehicle.splines[0])
            if getattr(self, '_all_free', False):
                self.free_ind[vehicle] = list(range(len(term_con)))
            conditions = self.define_variable('conT' + str(l), len(self.free_ind[vehicle]))
            cnt = 0
            self.term_con_len.append(len(term_con))
            for k, (spl, cond) in enumerate(term_con):
                if k in self.free_ind[vehicle]:
                    spline, condition = spl, conditions[cnt]
                    cnt += 1
                else:
                    spline, condition = spl, cond
                g = self.define_spline_variable('g' + str(k), 1, basis=spline.basis)[0]
                objective += definite_integral(g, self.t0, 1.)
                self.define_constraint(spline - condition - g, -inf, 0.)
                self.define_constraint(-spline + condition - g, -inf, 0.)
            # as in the reference (point2point.py:415-416) the loop over the derivative
            # conditions re-uses the LAST position spline and condition: it adds
            # len(term_con_der) copies of  spline(1) - condition = 0
            for _ in term_con_der:
                self.define_constraint(spline(1.) - condition, 0., 0.)
        self.define_objective(objective)


class FreeTPoint2point(Point2pointProblem):
    """Minimum-time point-to-point problem: the motion time T is a decision
    variable and the objective (reference point2point.py:269-374).

    The reference evaluates the initial conditions at ``t/T``; its parameter t
    is always 0 for this problem (point2point.py:300-306, the time axis resets
    every update) unless a multi-problem scheduler sets ``init_time``.  ``t/T``
    is not polynomial in the variable T, so the initial conditions are taken
    at 0 here and a non-zero ``init_time`` is refused.

    Note on the surveyed commit of the reference: ``FreeTPoint2point.construct``
    first runs ``Point2pointProblem.construct`` (point2point.py:53-62), which
    defines T as a *parameter* and hands that symbol to the vehicle and environment
    constraints, and only then defines the variable T (point2point.py:281-284) used by
    the objective; the parameter is never set.  This class implements the intended
    problem: every row depends on the variable T.
    """

    def __init__(self, fleet, environment, options):
        Point2pointProblem.__init__(self, fleet, environment, options)
        self.objective = 0.

    def define_time_symbols(self):
        self.T = self.define_variable('T', value=10.)
        self.t = self.define_parameter('t')
        self.t0 = 0.

    def construct(self):
        Point2pointProblem.construct(self)
        self.define_objective(self.T)
        self.define_constraint(-self.T, -inf, 0.)     # positive motion time
        self.define_init_constraints()
        self.define_terminal_constraints()

    def define_init_constraints(self):
        for vehicle in self.vehicles:
            init_con = vehicle.get_initial_constraints(vehicle.splines[0], self.T)
            for spline, condition in init_con:
                self.define_constraint(spline(0.) - condition, 0., 0.)

    def define_terminal_constraints(self):
        for vehicle in self.vehicles:
            term_con, term_con_der = vehicle.get_terminal_constraints(
                vehicle.splines[0])
            if self.options.get('no_term_con_der'):
                term_con_der = []
            for spline, condition in term_con + term_con_der:
                self.define_constraint(spline(1.) - condition, 0., 0.)

    def set_init_time(self, time):
        if time:
            raise NotImplementedError('FreeTPoint2point with a non-zero init_time: '
                                      't/T is not polynomial in T')
        self.init_time = time

    def set_parameters(self, current_time):
        parameters = Point2pointProblem.set_parameters(self, current_time)
        parameters[self]['t'] = 0.
        return parameters

    def horizon_time(self):
        return float(np.asarray(self.father.get_variables(self, 'T')).reshape(-1)[0])

    def store(self, current_time, update_time, sample_time):
        horizon_time = self.horizon_time()
        if horizon_time < sample_time:
            return
        for vehicle in self.vehicles:
            n_samp = int(round(horizon_time / sample_time, 6)) + 1
            time_axis = np.linspace(0., (n_samp - 1) * sample_time, n_samp)
            spline_segments = [self.father.get_variables(
                vehicle, 'splines_seg' + str(k)) for k in range(vehicle.n_seg)]
            vehicle.store(current_time, sample_time, spline_segments,
                          horizon_time, time_axis)

    def simulate(self, current_time, simulation_time, sample_time):
        horizon_time = self.horizon_time()
        if horizon_time < sample_time:
            return
        simulation_time = min(simulation_time, horizon_time)
        self.compute_partial_objective(current_time + simulation_time - self.start_time)
        Problem.simulate(self, current_time, simulation_time, sample_time)

    def stop_criterium(self, current_time, update_time):
        if self.horizon_time() < update_time:
            return True
        return Point2pointProblem.stop_criterium(self, current_time, update_time)

    def init_step(self, current_time, update_time):
        if (current_time - self.start_time) > 0:
            T = self.horizon_time()
            if T < 2 * update_time:      # almost arrived: shorten the update
                update_time = T - update_time
                target_time = T
            else:
                target_time = T - update_time
            # the piece from update_time on, re-expressed on an equidistant basis
            self.father.transform_primal_splines(
                lambda coeffs, basis, *_: shift_spline(
                    coeffs, update_time / target_time, basis))
            self.father.set_variables(target_time, self, 'T')

    def compute_partial_objective(self, current_time):
        self.objective = current_time

    def compute_objective(self):
        return self.objective
