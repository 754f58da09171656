"""Which variables does a coupling constraint share, and with whom?

The reference's ``DistributedProblem.interprete_constraints``
(omgtools/problems/distributedproblem.py:105-169) derives, from the constraints that couple the
agents of a distributed problem, three index structures per agent ("updater"):

    q_i [child][name]              the entries of agent i's own variables that appear in a coupling
    q_ij[other][child][name]       the entries of agent j's variables that agent i needs a copy of
    q_ji[other]                    = q_ij of the other agent seen from here

It reads them off CasADi's dependency information (``get_dependency`` on the MX expression).  Here
the constraints are polynomials (basics/poly.py) and the dependency of a row is simply the set of
symbols of its monomials, so the derivation works for any coupling -- formation constraints,
rendez-vous of terminal points, inter-vehicle hyperplanes -- and any mix of vehicle types.

``FormationPoint2point`` / ``RendezVous`` (problems/admm.py) use it to CHECK that the shared sets
their batched x-update is built for (all spline coefficients of the vehicle, resp. the free
terminal position) are the ones their coupling constraints imply.
"""
import collections as col

import numpy as np

from ..basics import poly as pl
from ..basics.poly import Poly


def variable_owners(agents):
    """{resolved symbol id: (agent, child label, entry name, flat index)} for the variables of
    ``agents`` = [[child, ...], ...] (the children that make up one agent: its vehicle, its
    problem, ...).  Entries flatten column-major like the NLP vectors (optilayer.py:225-243)."""
    owners = {}
    for a, children in enumerate(agents):
        for child in children:
            for name, mat in child._variables.items():
                flat = np.asarray(mat, dtype=object).reshape(-1, order='F')
                for k, sym in enumerate(flat):
                    owners[pl.resolve(sym.single_symbol())] = (a, child.label, name, k)
    return owners


def interprete_constraints(owners, constraints):
    """constraints: iterable of coupling constraints, each an array / list of Poly rows (or one
    Poly).  Returns (q_i, q_ij, q_ji): lists over the agents of ordered dicts as described in the
    module docstring; indices are sorted."""
    n_agents = 1 + max([o[0] for o in owners.values()] + [-1])
    q_i = [col.OrderedDict() for _ in range(n_agents)]
    q_ij = [col.OrderedDict() for _ in range(n_agents)]

    def add(dic, child, name, indices):
        cur = dic.setdefault(child, col.OrderedDict()).setdefault(name, [])
        cur[:] = sorted(set(cur) | set(indices))

    for con in constraints:
        rows = [con] if isinstance(con, Poly) else list(np.asarray(con, dtype=object).reshape(-1))
        dep = col.OrderedDict()                       # agent -> {(child, name): {indices}}
        for row in rows:
            if not isinstance(row, Poly):
                continue
            for sid in row.symbols():
                own = owners.get(pl.resolve(sid))
                if own is not None:
                    a, child, name, k = own
                    dep.setdefault(a, col.OrderedDict()).setdefault((child, name), set()).add(k)
        if len(dep) < 2:
            continue                                  # not a coupling constraint
        for a, mine in dep.items():
            for (child, name), idx in mine.items():
                add(q_i[a], child, name, idx)
            for b, theirs in dep.items():
                if b == a:
                    continue
                for (child, name), idx in theirs.items():
                    add(q_ij[a].setdefault(b, col.OrderedDict()), child, name, idx)
    for a in range(n_agents):                         # neighbours in index order (distributedproblem.py:163-169)
        q_ij[a] = col.OrderedDict(sorted(q_ij[a].items()))
    q_ji = [col.OrderedDict((b, q_ij[b][a]) for b in range(n_agents) if a in q_ij[b]) for a in range(n_agents)]
    return q_i, q_ij, q_ji
