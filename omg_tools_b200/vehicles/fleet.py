"""Fleet: the vehicles of a multi-agent problem, who talks to whom, and the formation they keep.

API of the reference's ``omgtools/vehicles/fleet.py`` (OMG-tools, Copyright (C) 2016 Ruben Van
Parys & Tim Mercy, KU Leuven, LGPL v3) because the problem classes address it by these names --
``vehicles``, ``N``, ``get_neighbors``, ``set_configuration``, ``get_rel_config``,
``configuration``, each vehicle's ``rel_pos_c`` -- re-implemented here on index arrays:
the neighbour table is an integer matrix (what the batched ADMM runner and the NCCL exchange
consume directly, problems/admm.py), the configuration a dense array plus the per-vehicle
dictionaries the reference exposes (fleet.py:46-98)."""
import numpy as np

from .vehicle import Vehicle

TOPOLOGIES = {
    # ring: successor first, then predecessor (the order fixes the layout of z_ij / l_ij)
    'circular': lambda n: np.stack([(np.arange(n) + 1) % n, (np.arange(n) - 1) % n], axis=1),
    'full': lambda n: np.array([[k for k in range(n) if k != i] for i in range(n)], dtype=int).reshape(n, -1),
    # open chain (not in the reference): the end vehicles have ONE neighbour.  -1 marks a missing
    # neighbour; the batched ADMM pads such slots with the agent itself (problems/admm.py)
    'line': lambda n: np.stack([np.where(np.arange(n) + 1 < n, np.arange(n) + 1, -1),
                                np.arange(n) - 1], axis=1),
}


def get_fleet_vehicles(var):
    """Accepts a Fleet, a Vehicle, a list of vehicles or a one-element list holding a Fleet;
    returns (fleet, vehicles)."""
    if isinstance(var, Vehicle):
        var = [var]
    if isinstance(var, list) and var and isinstance(var[0], Fleet):
        var = var[0]
    fleet = var if isinstance(var, Fleet) else Fleet(list(var))
    return fleet, fleet.vehicles


class Fleet(object):

    def __init__(self, vehicles=None, interconnection='circular'):
        if vehicles is None:
            vehicles = []
        self.vehicles = list(vehicles) if isinstance(vehicles, (list, tuple)) else [vehicles]
        self.interconnection = interconnection
        self.set_neighbors()

    # ---- topology ----------------------------------------------------------------------
    def set_neighbors(self):
        self.N = len(self.vehicles)
        if isinstance(self.interconnection, str):
            if self.interconnection not in TOPOLOGIES:
                raise ValueError('Interconnection type ' + str(self.interconnection) + ' not understood.')
            self.nghb_index = TOPOLOGIES[self.interconnection](self.N) if self.N else np.zeros((0, 0), int)
        else:
            # explicit adjacency lists [[j, ...], ...] with different lengths: padded with -1
            lists = [list(row) for row in self.interconnection]
            width = max([len(row) for row in lists] + [0])
            self.nghb_index = np.array([row + [-1] * (width - len(row)) for row in lists], dtype=int).reshape(self.N, width)
        self.nghb_list = {veh: [self.vehicles[k] for k in self.nghb_index[i] if k >= 0]
                          for i, veh in enumerate(self.vehicles)}

    def get_neighbors(self, vehicle):
        return self.nghb_list[vehicle]

    # ---- formation ----------------------------------------------------------------------
    def set_configuration(self, configuration, orientation=0.):
        """configuration[i]: the place of vehicle i relative to the formation centre -- a list of
        coordinates (2-D ones are turned by -orientation) or a dict {spline index: offset}."""
        if len(configuration) != self.N:
            raise ValueError('You should provide configuration info for each vehicle.')
        turn = np.array([[np.cos(orientation), np.sin(orientation)],
                         [-np.sin(orientation), np.cos(orientation)]])
        self.configuration = {}
        for veh, entry in zip(self.vehicles, configuration):
            if isinstance(entry, dict):
                self.configuration[veh] = dict(entry)
                continue
            coords = np.asarray(entry, dtype=float)
            if coords.size == 2:
                coords = turn.dot(coords)
            self.configuration[veh] = dict(enumerate(coords.tolist()))
        self.set_rel_pos_c()
        keys = {veh: sorted(self.configuration[veh]) for veh in self.vehicles}
        self.rel_config = {}
        for veh in self.vehicles:
            mine = np.array([self.configuration[veh][k] for k in keys[veh]])
            self.rel_config[veh] = {}
            for other in self.nghb_list[veh]:
                if len(keys[other]) != len(keys[veh]):
                    raise ValueError('All vehicles should have same number of variables for which '
                                     'the configuration is imposed.')
                theirs = np.array([self.configuration[other][k] for k in keys[other]])
                self.rel_config[veh][other] = (mine - theirs).tolist()

    def set_rel_pos_c(self):
        """Vector from a vehicle to the formation centre (a parameter of the formation rows)."""
        for veh in self.vehicles:
            veh.rel_pos_c = [-self.configuration[veh][k] for k in sorted(self.configuration[veh])]

    def get_rel_config(self, vehicle):
        return self.rel_config[vehicle]

    # ---- boundary conditions, vehicle by vehicle ---------------------------------------------
    def set_initial_conditions(self, states, inputs=None):
        inputs = [None] * len(states) if inputs is None else inputs
        for veh, state, inp in zip(self.vehicles, states, inputs):
            veh.set_initial_conditions(state, inp)

    def set_terminal_conditions(self, conditions):
        for veh, cond in zip(self.vehicles, conditions):
            veh.set_terminal_conditions(cond)
