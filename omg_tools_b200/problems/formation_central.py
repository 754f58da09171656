"""Formation point-to-point as ONE coupled NLP: all vehicles of the fleet, their collision
constraints, and the formation constraints centre_i(tau) = centre_j(tau) between neighbours,
hard (equality rows) or soft (slack splines with an L1 penalty and an optional maximum
deviation).  Reference ``omgtools/problems/formation_central.py`` (options 30-34, construct
36-78, parameters 84-90); it is the problem the ADMM of problems/admm.py distributes, and
tests/test_admm.py uses a coupled solve as the fixed point the ADMM iteration must reach.

As in the reference the terminal slack splines 'g0', 'g1' are defined once per vehicle
under the same name (point2point.py:160-163) and are therefore ONE pair of variables shared
by the whole fleet (basics/optilayer.py: entries are identified by name)."""
from .point2point import FixedTPoint2point
from ..basics.optilayer import inf
from ..basics.spline_extra import definite_integral


class FormationPoint2pointCentral(FixedTPoint2point):

    def __init__(self, fleet, environment, options=None):
        FixedTPoint2point.__init__(self, fleet, environment, options)

    def set_default_options(self):
        FixedTPoint2point.set_default_options(self)
        self.options['soft_formation'] = False
        self.options['soft_formation_weight'] = 10
        self.options['max_formation_deviation'] = inf

    def construct(self):
        config = self.fleet.configuration
        rel_pos_c = {}
        for veh in self.vehicles:
            ind_veh = sorted(config[veh].keys())
            rel_pos_c[veh] = veh.define_parameter('rel_pos_c', len(ind_veh))
        FixedTPoint2point.construct(self)
        centra = {}
        for veh in self.vehicles:
            splines = self.father.get_variables(veh, 'splines_seg0', symbolic=True)
            centra[veh] = veh.get_fleet_center(splines, rel_pos_c[veh], substitute=False)
        # every neighbour pair once; a circular fleet would close the loop with
        # redundant rows, so the last two vehicles add none (formation_central.py:55-57)
        couples = {veh: [] for veh in self.vehicles}
        for veh in self.vehicles:
            for nghb in self.fleet.get_neighbors(veh):
                if veh not in couples[nghb] and nghb not in couples[veh]:
                    couples[veh].append(nghb)
        if self.fleet.interconnection == 'circular':
            couples.pop(self.vehicles[-1], None)
            couples.pop(self.vehicles[-2], None)
        t = self.define_symbol('t')
        T = self.define_symbol('T')
        for veh, nghbs in couples.items():
            ind_veh = sorted(config[veh].keys())
            for nghb in nghbs:
                ind_nghb = sorted(config[nghb].keys())
                for ind_v, ind_n in zip(ind_veh, ind_nghb):
                    diff = centra[veh][ind_v] - centra[nghb][ind_n]
                    if self.options['soft_formation']:
                        weight = self.options['soft_formation_weight']
                        eps = self.define_spline_variable(
                            'eps_form_' + str(ind_v) + str(ind_n), basis=veh.basis)[0]
                        self.define_objective(weight * definite_integral(eps, t / T, 1.))
                        self.define_constraint(diff - eps, -inf, 0.)
                        self.define_constraint(-diff - eps, -inf, 0.)
                        if self.options['max_formation_deviation'] != inf:
                            max_dev = abs(self.options['max_formation_deviation'])
                            self.define_constraint(eps, -max_dev, max_dev)
                    else:
                        self.define_constraint(diff, 0., 0.)

    def set_parameters(self, current_time):
        parameters = FixedTPoint2point.set_parameters(self, current_time)
        for veh in self.vehicles:
            if veh not in parameters:
                parameters[veh] = {}
            parameters[veh].update({'rel_pos_c': veh.rel_pos_c})
        return parameters
