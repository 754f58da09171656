"""Problem base class: options, model construction and the per-MPC-step
``solve()`` -- the hot path of this framework.

Interface of the reference's ``omgtools/problems/problem.py`` (children order
45-48, options 54-74, init 85-91, solve 103-136, predict 138-163,
reset_init_guess 165-181).  The single line the reference spends its time in,
``self.problem(x0=var, p=par, lbg=lb, ubg=ub)`` (problem.py:113), is kept
verbatim: ``self.problem`` is the B200 solver object created by
``OptiFather.construct_problem`` / ``create_nlp``.
"""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/problems/problem.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
from __future__ import print_function

import time

import numpy as np

from ..basics.optilayer import OptiFather, OptiChild
from ..vehicles.fleet import get_fleet_vehicles


class Problem(OptiChild):

    def __init__(self, fleet, environment, options=None, label='problem'):
        options = options or {}
        OptiChild.__init__(self, label)
        self.fleet, self.vehicles = get_fleet_vehicles(fleet)
        self.environment = environment
        self.set_default_options()
        self.set_options(options)
        self.iteration = 0
        self.update_times = []
        children = [vehicle for vehicle in self.vehicles]
        children += [obstacle for obstacle in self.environment.obstacles]
        children += [self, self.environment]
        self.father = OptiFather(children)

    # ------------------------------------------------------------------
    # options
    # ------------------------------------------------------------------

    def set_default_options(self):
        self.options = {'verbose': 2}
        self.options['solver'] = 'b200'
        # the reference's IPOPT settings (problem.py:57-60) are the defaults
        # of the B200 interior-point solver as well
        b200_options = {'tol': 1e-3, 'warm_start_init_point': 'yes',
                        'print_level': 0}
        self.options['solver_options'] = {'b200': b200_options}
        self.options['codegen'] = {'build': None, 'flags': '-O0'}

    def set_options(self, options):
        if 'solver_options' in options:
            for key, value in options['solver_options'].items():
                if key not in self.options['solver_options']:
                    self.options['solver_options'][key] = {}
                self.options['solver_options'][key].update(value)
        if 'codegen' in options:
            self.options['codegen'].update(options['codegen'])
        for key in options:
            if key not in ['solver_options', 'codegen']:
                self.options[key] = options[key]

    # ------------------------------------------------------------------
    # create problem
    # ------------------------------------------------------------------

    def construct(self):
        self.environment.init()
        for vehicle in self.vehicles:
            vehicle.init()

    def init(self, problem=None):
        self.father.reset()
        self.construct()
        self.problem, buildtime = self.father.construct_problem(
            self.options, problem=problem)
        self.father.init_transformations(self.init_primal_transform,
                                         self.init_dual_transform)
        return buildtime

    # ------------------------------------------------------------------
    # deploying
    # ------------------------------------------------------------------

    def reinitialize(self, father=None):
        if father is None:
            father = self.father
        father.init_variables()
        father.init_parameters()

    def solve(self, current_time, update_time):
        current_time -= self.start_time
        self.init_step(current_time, update_time)
        var = self.father.get_variables()
        par = self.father.set_parameters(current_time)
        lb, ub = self.father.update_bounds(current_time)
        t0 = time.time()
        result = self.problem(x0=var, p=par, lbg=lb, ubg=ub)
        t1 = time.time()
        t_upd = t1 - t0
        self.father.set_variables(result['x'])
        self.father.set_dual_variables(result['lam_g'])
        stats = self.problem.stats()
        if stats['return_status'] != 'Solve_Succeeded':
            if stats['return_status'] == 'Maximum_CpuTime_Exceeded':
                if current_time != 0.0:
                    print('Maximum solving time exceeded, resetting initial guess')
                    self.reset_init_guess()
                    print(stats['return_status'])
            else:
                print(stats['return_status'])
        if self.options['verbose'] >= 2:
            self.iteration += 1
            if (self.iteration - 1) % 20 == 0:
                print("----|------------|------------")
                print("%3s | %10s | %10s " % ("It", "t upd", "time"))
                print("----|------------|------------")
            print("%3d | %.4e | %.4e " % (self.iteration, t_upd, current_time))
        self.update_times.append(t_upd)

    def predict(self, current_time, predict_time, sample_time, states=None,
                inputs=None, dinputs=None, delay=0, enforce_states=False,
                enforce_inputs=False):
        """Reference problem.py:138-163: per-vehicle prediction; a measured state is enforced on
        the first iteration (current_time == start_time)."""
        nv = len(self.vehicles)

        def per_vehicle(arg):
            if arg is None:
                return [None] * nv
            if nv == 1 and (not isinstance(arg, list) or isinstance(arg[0], float)):
                return [arg]
            return arg
        states, inputs, dinputs = per_vehicle(states), per_vehicle(inputs), per_vehicle(dinputs)
        if current_time == self.start_time:
            enforce_states = True
        for k, vehicle in enumerate(self.vehicles):
            vehicle.predict(current_time, predict_time, sample_time, states[k], inputs[k],
                            dinputs[k], delay, enforce_states, enforce_inputs)

    def reset_init_guess(self, init_guess=None):
        if init_guess is None:
            init_guess = [vehicle.get_init_spline_value()
                          for vehicle in self.vehicles]
        for vehicle, guess in zip(self.vehicles, init_guess):
            guess = guess if isinstance(guess, list) else [guess]
            if len(guess) != vehicle.n_seg:
                raise ValueError('Each spline segment of the vehicle should '
                                 'receive an initial guess.')
            for l in range(vehicle.n_seg):
                if guess[l].shape[1] != vehicle.n_spl:
                    raise ValueError('Each vehicle spline should receive an '
                                     'initial guess.')
                self.father.set_variables(guess[l], child=vehicle,
                                          name='splines_seg' + str(l))

    # ------------------------------------------------------------------
    # simulation
    # ------------------------------------------------------------------

    def simulate(self, current_time, simulation_time, sample_time):
        for vehicle in self.vehicles:
            vehicle.simulate(simulation_time, sample_time)
        self.environment.simulate(simulation_time, sample_time)

    # ------------------------------------------------------------------
    # methods encouraged to override
    # ------------------------------------------------------------------

    def init_step(self, current_time, update_time):
        pass

    def final(self):
        pass

    def initialize(self, current_time):
        pass

    def init_primal_transform(self, basis):
        return None

    def init_dual_transform(self, basis):
        return None

    def set_parameters(self, time):
        return {self: {}}

    def update(self, current_time, update_time, sample_time):
        raise NotImplementedError('Please implement this method!')

    def store(self, current_time, update_time, sample_time):
        raise NotImplementedError('Please implement this method!')

    def stop_criterium(self, current_time, update_time):
        raise NotImplementedError('Please implement this method!')
