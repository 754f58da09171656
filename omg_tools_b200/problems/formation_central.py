"""The formation problem as ONE coupled NLP -- the fixed point the distributed ADMM iteration of
problems/admm.py has to reach, which is what it is here for: tests/test_admm.py solves it with
the oracle and checks that the ADMM iterates converge to it.  (As a user-facing problem class it
is outside this build's scope)

Model of the reference's ``omgtools/problems/formation_central.py:30-90`` (OMG-tools, Copyright (C)
2016 Ruben Van Parys & Tim Mercy, KU Leuven, LGPL v3): every vehicle gets a parameter
``rel_pos_c`` (its offset to the formation centre); neighbouring vehicles must agree on the
centre, ``centre_i(tau) - centre_j(tau) = 0`` coefficient by coefficient -- hard, or softened by a
slack spline with an L1 penalty and an optional cap.  The option names are the reference's.

Entries are identified by NAME across the children (basics/optilayer.py): the terminal slacks
'g0', 'g1' that every vehicle's terminal constraints define (point2point.py:160-163) are one
shared pair, exactly as in the reference."""
from .point2point import FixedTPoint2point
from ..basics.optilayer import inf
from ..basics.spline_extra import definite_integral


def coupled_pairs(fleet):
    """Each neighbour relation once, as (vehicle, neighbour).  On a ring the last two vehicles
    contribute none: their relations would close the loop with linearly dependent rows."""
    seen, pairs = set(), []
    skip = set(fleet.vehicles[-2:]) if fleet.interconnection == 'circular' else set()
    for veh in fleet.vehicles:
        for other in fleet.get_neighbors(veh):
            key = frozenset((veh, other))
            if key in seen:
                continue
            seen.add(key)
            if veh not in skip:
                pairs.append((veh, other))
    return pairs


class FormationPoint2pointCentral(FixedTPoint2point):

    def set_default_options(self):
        FixedTPoint2point.set_default_options(self)
        self.options.update({'soft_formation': False, 'soft_formation_weight': 10,
                             'max_formation_deviation': inf})

    def construct(self):
        offsets = {veh: veh.define_parameter('rel_pos_c', len(self.fleet.configuration[veh]))
                   for veh in self.vehicles}
        FixedTPoint2point.construct(self)
        centre = {veh: veh.get_fleet_center(self.father.get_variables(veh, 'splines_seg0', symbolic=True),
                                            offsets[veh], substitute=False)
                  for veh in self.vehicles}
        self._t_rel = self.define_symbol('t') / self.define_symbol('T')
        for veh, other in coupled_pairs(self.fleet):
            mine, theirs = sorted(self.fleet.configuration[veh]), sorted(self.fleet.configuration[other])
            for a, b in zip(mine, theirs):
                self._couple(centre[veh][a] - centre[other][b], veh.basis, '%d%d' % (a, b))

    def _couple(self, gap, basis, tag):
        if not self.options['soft_formation']:
            self.define_constraint(gap, 0., 0.)
            return
        slack = self.define_spline_variable('eps_form_' + tag, basis=basis)[0]
        self.define_objective(self.options['soft_formation_weight'] * definite_integral(slack, self._t_rel, 1.))
        self.define_constraint(gap - slack, -inf, 0.)
        self.define_constraint(-gap - slack, -inf, 0.)
        cap = self.options['max_formation_deviation']
        if cap != inf:
            self.define_constraint(slack, -abs(cap), abs(cap))

    def set_parameters(self, current_time):
        parameters = FixedTPoint2point.set_parameters(self, current_time)
        for veh in self.vehicles:
            parameters.setdefault(veh, {})['rel_pos_c'] = veh.rel_pos_c
        return parameters
