"""Environment: room + obstacles; creates the separating-hyperplane spline
variables a(t), b(t) per (vehicle shape, obstacle) and the ||a||^2 <= 1 rows
(reference ``omgtools/environment/environment.py``: room handling 32-61,
add_obstacle 75-92, define_collision_constraints 102-146, init 182-184)."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/environment/environment.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import warnings

import numpy as np

from ..basics.optilayer import OptiChild, inf
from ..basics.spline import BSplineBasis, BSpline
from .obstacle import Obstacle


class Environment(OptiChild):

    def __init__(self, room, obstacles=None):
        obstacles = obstacles or []
        OptiChild.__init__(self, 'environment')
        self.room = room if isinstance(room, list) else [room]
        self.n_dim = self.room[0]['shape'].n_dim
        for room in self.room:
            if room['shape'].n_dim != self.n_dim:
                raise ValueError('You try to combine rooms of different dimensions,' +
                                 ' which is invalid')
            if 'position' not in room:
                room['position'] = [0. for k in range(self.n_dim)]
            if 'orientation' not in room:
                room['orientation'] = 0. if self.n_dim == 2 else [0., 0., 0.]
            if 'draw' not in room:
                room['draw'] = False
        self.obstacles, self.n_obs = [], 0
        for obstacle in obstacles:
            self.add_obstacle(obstacle)

    def copy(self):
        obstacles = [Obstacle(o.initial, o.shape, o.simulation, o.options)
                     for o in self.obstacles]
        return Environment(self.room, obstacles)

    def add_obstacle(self, obstacle):
        if isinstance(obstacle, list):
            for obst in obstacle:
                self.add_obstacle(obst)
            return
        if obstacle.n_dim == 2 and self.n_dim == 3:
            warnings.warn('You are combining a 2D obstacle with a 3D '
                          'environment: it is extended infinitely in z.')
        if obstacle.n_dim == 3 and self.n_dim == 2:
            raise ValueError('Not possible to combine ' +
                             str(obstacle.n_dim) + 'D obstacle with ' +
                             str(self.n_dim) + 'D environment.')
        self.obstacles.append(obstacle)
        self.n_obs += 1

    def define_collision_constraints(self, vehicle, splines, horizon_times):
        if vehicle.n_dim != self.n_dim:
            raise ValueError('Not possible to combine ' +
                             str(vehicle.n_dim) + 'D vehicle with ' +
                             str(self.n_dim) + 'D environment.')
        horizon_times = horizon_times if isinstance(horizon_times, list) \
            else [horizon_times]
        degree = 1
        knots = np.r_[np.zeros(degree),
                      vehicle.knots[vehicle.degree:-vehicle.degree],
                      np.ones(degree)]
        basis = BSplineBasis(knots, degree)
        for idx in range(vehicle.n_seg):
            room = self.room[idx]
            hyp_veh, hyp_obs = {}, {}
            obs_to_add = room['obstacles'] if 'obstacles' in room \
                else self.obstacles
            for k, shape in enumerate(vehicle.shapes):
                hyp_veh[shape] = []
                for l, obstacle in enumerate(obs_to_add):
                    obstacle.init(horizon_times=horizon_times[:idx + 1])
                    if obstacle.options['avoid']:
                        if obstacle not in hyp_obs:
                            hyp_obs[obstacle] = []
                        tag = '_' + vehicle.label + '_' + 'seg' + str(idx) + \
                            '_' + str(k) + str(l)
                        a = self.define_spline_variable(
                            'a' + tag, obstacle.n_dim, basis=basis)
                        b = self.define_spline_variable(
                            'b' + tag, 1, basis=basis)[0]
                        self.define_constraint(
                            sum([a[p] * a[p] for p in range(obstacle.n_dim)]) - 1,
                            -inf, 0.)
                        if self.n_dim == 3 and obstacle.n_dim == 2:
                            a2 = [a[0], a[1], BSpline(basis, np.zeros(len(basis)))]
                            hyp_veh[shape].append({'a': a2, 'b': b})
                        else:
                            hyp_veh[shape].append({'a': a, 'b': b})
                        hyp_obs[obstacle].append({'a': a, 'b': b})
                        obstacle.define_collision_constraints(hyp_obs[obstacle])
            vehicle.define_collision_constraints(hyp_veh, room, splines[idx],
                                                 horizon_times[idx])

    def define_intervehicle_collision_constraints(self, vehicles, horizon_times):
        """Separating hyperplanes between every pair of vehicles (one per pair of shapes):
        a, b on the degree-1 basis over the union of both vehicles' knots; vehicle k sees
        (a, b), vehicle l the mirrored (-a, -b) (reference environment.py:148-175)."""
        horizon_times = horizon_times if isinstance(horizon_times, list) \
            else [horizon_times] * vehicles[0].n_seg
        for idx in range(vehicles[0].n_seg):
            hyp_veh = {veh: {sh: [] for sh in veh.shapes} for veh in vehicles}
            for k in range(len(vehicles)):
                for l in range(k + 1, len(vehicles)):
                    veh1, veh2 = vehicles[k], vehicles[l]
                    if veh1.n_dim != veh2.n_dim:
                        raise ValueError('Not possible to combine ' + str(veh1.n_dim) +
                                         'D and ' + str(veh2.n_dim) + 'D vehicle.')
                    degree = 1
                    knots = np.r_[np.zeros(degree),
                                  np.union1d(veh1.knots[veh1.degree:-veh1.degree],
                                             veh2.knots[veh2.degree:-veh2.degree]),
                                  np.ones(degree)]
                    basis = BSplineBasis(knots, degree)
                    for kk, shape1 in enumerate(veh1.shapes):
                        for ll, shape2 in enumerate(veh2.shapes):
                            tag = '_' + veh1.label + '_' + 'seg' + str(idx) + '_' + str(kk) + \
                                '_' + veh2.label + '_' + str(ll)
                            a = self.define_spline_variable('a' + tag, self.n_dim, basis=basis)
                            b = self.define_spline_variable('b' + tag, 1, basis=basis)[0]
                            self.define_constraint(
                                sum([a[p] * a[p] for p in range(self.n_dim)]) - 1, -inf, 0.)
                            hyp_veh[veh1][shape1].append({'a': a, 'b': b})
                            hyp_veh[veh2][shape2].append({'a': [-a_i for a_i in a], 'b': -b})
            for vehicle in vehicles:
                vehicle.define_collision_constraints(hyp_veh[vehicle], self.room[idx],
                                                     vehicle.splines[idx], horizon_times[idx])

    def init(self, horizon_times=None):
        for obstacle in self.obstacles:
            obstacle.init(horizon_times=horizon_times)

    def simulate(self, simulation_time, sample_time):
        for obstacle in self.obstacles:
            obstacle.simulate(simulation_time, sample_time)
