"""ADMM for formation point-to-point problems, batched over the agents.

Reference: ``omgtools/problems/admm.py`` (ADMM updater: x-update NLP 63-115,
closed-form z-update 117-168, lambda-update 248-266, residuals 268-307,
communicate 468-475, knot shift 477-491, dual_update 584-628),
``dualmethod.py`` (init_iter / max_iter_per_update, 200-224) and
``formation.py`` (coupling constraints 33-72).

In the reference the N agents are N Python objects updated in a serial loop and
"communication" is attribute copying.  Here all agents share ONE NLP structure
(identical vehicles, equal neighbour count), so

  * the x-update of all local agents is one ``omg_solve_batch`` call
    (the parameters z_i, z_ji, l_i, l_ji, rho make the instances differ);
  * the z-update, lambda-update and residuals of all local agents are one
    ``omg_admm_zl_update`` kernel launch.  The z-update is the reference's
    equality-constrained QP  z = v + A^T (A A^T)^-1 (b - A v),  v = x + l/rho,
    in first-knot-shifted coordinates (admm.py:143-163); A is constant, so the
    projector P = I - A^T (A A^T)^-1 A and c_i = A^T (A A^T)^-1 b_i are formed
    once on the host;
  * agents are sharded contiguously over the ranks; only the neighbour
    exchange crosses GPUs (NCCL send/recv of x_i, then z_ij/l_ij) and the three
    squared residuals are all-reduced (admm.py:597-605).
"""
from __future__ import print_function

import numpy as np

from ..basics.optilayer import OptiChild, OptiFather
from ..basics.poly import Poly
from ..basics import poly as pl
from ..basics.spline import BSpline
from ..basics.spline_extra import shift_knot1_fwd, shiftfirstknot_T, shiftoverknot_T
from .point2point import Point2point


class ADMMUpdater(OptiChild):
    """Holds the consensus parameters and the augmented-Lagrangian objective of
    one agent's x-update (reference admm.py:63-115)."""

    def __init__(self):
        OptiChild.__init__(self, 'admm')

    def construct(self, vehicle, p2p, n_nghb, ama=False, shared=None):
        """shared = None: the vehicle's spline coefficients (formation); otherwise a flat
        array of symbols that are not splines (the free terminal conditions of a
        RendezVous): no first-knot transformation."""
        if shared is not None:
            return self._construct_plain(np.asarray(shared).reshape(-1), n_nghb, ama)
        L, ns = len(vehicle.basis), vehicle.n_spl
        nsh = L * ns
        z_i = self.define_parameter('z_i', nsh)
        z_ji = self.define_parameter('z_ji', nsh * n_nghb)
        l_i = self.define_parameter('l_i', nsh)
        l_ji = self.define_parameter('l_ji', nsh * n_nghb)
        rho = self.define_parameter('rho')
        t = self.define_symbol('t')
        T = self.define_symbol('T')
        t0 = t / T
        basis = vehicle.basis
        x_i = vehicle._variables['splines_seg0']          # (L, ns) symbols

        def fwd(vec):            # future piece of every spline of the block
            out = []
            for k in range(ns):
                out.append(shift_knot1_fwd(vec[k * L:(k + 1) * L], basis, t0))
            return np.concatenate(out)

        x = fwd(x_i.reshape(-1, order='F'))
        obj = Poly()

        def add(z, l):
            nonlocal obj
            z, l = fwd(z), fwd(l)
            for k in range(nsh):
                d = x[k] - z[k]
                obj = obj + l[k] * d
                if not ama:
                    obj = obj + 0.5 * rho * d * d

        add(np.asarray(z_i), np.asarray(l_i))
        for j in range(n_nghb):
            add(np.asarray(z_ji)[j * nsh:(j + 1) * nsh], np.asarray(l_ji)[j * nsh:(j + 1) * nsh])
        self.define_objective(obj)


def _plain_objective(updater, x, n_nghb, ama):
    nsh = len(x)
    z_i = updater.define_parameter('z_i', nsh)
    z_ji = updater.define_parameter('z_ji', nsh * n_nghb)
    l_i = updater.define_parameter('l_i', nsh)
    l_ji = updater.define_parameter('l_ji', nsh * n_nghb)
    rho = updater.define_parameter('rho')
    obj = Poly()
    blocks = [(np.asarray(z_i), np.asarray(l_i))]
    blocks += [(np.asarray(z_ji)[j * nsh:(j + 1) * nsh], np.asarray(l_ji)[j * nsh:(j + 1) * nsh])
               for j in range(n_nghb)]
    for z, l in blocks:
        for k in range(nsh):
            d = x[k] - z[k]
            obj = obj + l[k] * d
            if not ama:
                obj = obj + 0.5 * rho * d * d
    updater.define_objective(obj)


ADMMUpdater._construct_plain = lambda self, x, n_nghb, ama: _plain_objective(self, x, n_nghb, ama)


def _linear_rows(rows, var_syms, par_vals):
    """rows: Poly linear in var_syms with parameter-only remainder.
    Returns (A, b) with A z = b."""
    index = {s: k for k, s in enumerate(var_syms)}
    A = np.zeros((len(rows), len(var_syms)))
    b = np.zeros(len(rows))
    for i, r in enumerate(rows):
        for mono, c in r.t.items():
            vs = [s for s in mono if pl.resolve(s) in index]
            ps = [s for s in mono if pl.resolve(s) not in index]
            if len(vs) > 1:
                raise ValueError('coupling constraint is not linear')
            val = c
            for s in ps:
                val *= par_vals[pl.resolve(s)]
            if vs:
                A[i, index[pl.resolve(vs[0])]] += val
            else:
                b[i] -= val
    return A, b


class FormationPoint2point(object):
    """Fleet of identical vehicles keeping a formation while moving point to
    point (reference formation.py:26-72), solved by ADMM on the GPU(s)."""

    def __init__(self, fleet, environment, options=None, rank=0, world=1, group=None):
        self.fleet = fleet
        self.vehicles = fleet.vehicles
        self.N = len(self.vehicles)
        self.environment = environment
        self.rank, self.world, self.group = rank, world, group
        self.options = {'verbose': 0, 'rho': 2., 'init_iter': 5, 'max_iter_per_update': 1,
                        'AMA': False, 'horizon_time': 10.,
                        # 'AMA': alternating minimisation (no quadratic penalty in the x-update,
                        # admm.py:97-104); fast ADMM / fast AMA (admm.py:33-37, 510-554)
                        'nesterov_acceleration': False, 'nesterov_reset': False, 'eta': 0.999}
        self.options.update(options or {})
        self.iteration = 0
        self.residuals = {'primal': [], 'dual': [], 'combined': []}
        self.index = {v: k for k, v in enumerate(self.vehicles)}
        # neighbour table: one row per agent, as wide as the largest neighbourhood.  Agents with
        # fewer neighbours (an open chain, an arbitrary graph) get the missing slots filled with
        # THEMSELVES: a copy z_ii that must agree with z_i is a redundant consensus constraint --
        # it leaves the fixed point untouched and keeps ONE NLP structure and ONE projector for the
        # whole batch (the reference builds per-agent problems instead, admm.py:63-168).
        table = np.array(fleet.nghb_index, dtype=np.int64).reshape(self.N, -1)
        own = np.arange(self.N)[:, None]
        self.real_nghb = table >= 0
        self.nghb = np.where(table >= 0, table, own)
        self.n_nghb = self.nghb.shape[1]
        # slot of agent i in neighbour j's list (who holds z_ij for me); a self slot maps to itself
        self.back = np.zeros_like(self.nghb)
        for i in range(self.N):
            for k, j in enumerate(self.nghb[i]):
                self.back[i, k] = k if j == i else list(self.nghb[j]).index(i)

    # ------------------------------------------------------------------
    def init(self, build_solver=True):
        veh = self.vehicles[0]
        p2p_opts = {k: v for k, v in self.options.items()
                    if k in ('horizon_time', 'solver', 'solver_options', 'verbose',
                             'hard_term_con', 'no_term_con_der')}
        self.p2p = Point2point(veh, self.environment.copy(), p2p_opts, freeT=False)
        self.updater = ADMMUpdater()
        env = self.p2p.environment
        children = [veh, self.p2p, env, self.updater] + list(env.obstacles)
        father = OptiFather(children)
        father.reset()
        # formation.py:33-40: rel_pos_c parameter and the fleet centre first
        self.rel_pos_c = veh.define_parameter('rel_pos_c', veh.n_dim)
        self.p2p.father = father
        self.p2p.construct()
        self.updater.construct(veh, self.p2p, self.n_nghb, self.options['AMA'])
        self.father = father
        if build_solver:
            self.solver, _ = father.construct_problem(self.p2p.options)
        else:
            father.translate_symbols()
            father.construct_variables()
            father.construct_parameters()
            rows, lb, ub = father.construct_constraints()
            from ..basics.lowering import lower
            father.tables = lower(father._var_ids, father._par_ids, rows,
                                  father.construct_objective(), lb, ub, father.order_hint())
            father.init_variables()
            father.init_parameters()
            self.solver = None
        father.init_transformations(self.p2p.init_primal_transform,
                                    self.p2p.init_dual_transform)
        self.tb = father.tables
        self.basis = veh.basis
        self.L, self.ns = len(veh.basis), veh.n_spl
        self.nsh = self.L * self.ns
        self.knot_time = self.p2p.knot_time
        self.T = self.options['horizon_time']
        ent = father._var_struct.entries[(veh.label, 'splines_seg0')]
        self.x_off = ent[0]
        self.par_off = {k: v[0] for k, v in father._par_struct.entries.items()}
        self.veh_label, self.p2p_label, self.upd_label = veh.label, self.p2p.label, self.updater.label
        self._build_consensus_projector()
        self._derive_shared_sets()
        self._init_agent_data()

    def _derive_shared_sets(self):
        """q_i / q_ij / q_ji from the coupling constraints of the whole fleet (the reference's
        interprete_constraints, distributedproblem.py:105-169; here problems/distributed.py), and
        the check that they are what the batched x-update is built for: every agent shares ALL
        coefficients of its vehicle splines, and needs copies from exactly its fleet neighbours."""
        from .distributed import interprete_constraints
        L, ns = self.L, self.ns
        syms = [pl.sym_array('q%d' % i, 'var', L * ns, 1)[:, 0] for i in range(self.N)]
        owners = {s.single_symbol(): (i, self.vehicles[i].label, 'splines_seg0', k)
                  for i in range(self.N) for k, s in enumerate(syms[i])}
        cons = []
        for i in range(self.N):
            for k_slot, j in enumerate(self.nghb[i]):
                if not self.real_nghb[i, k_slot]:
                    continue
                for k in range(ns):
                    ci = BSpline(self.basis, syms[i][k * L:(k + 1) * L]) + float(self.vehicles[i].rel_pos_c[k])
                    cj = BSpline(self.basis, syms[j][k * L:(k + 1) * L]) + float(self.vehicles[j].rel_pos_c[k])
                    cons.append((ci - cj).coeffs)
        self.q_i, self.q_ij, self.q_ji = interprete_constraints(owners, cons)
        for i in range(self.N):
            want = sorted(set(int(j) for k, j in enumerate(self.nghb[i]) if self.real_nghb[i, k]))
            got_i = self.q_i[i].get(self.vehicles[i].label, {}).get('splines_seg0', [])
            if got_i != list(range(L * ns)) or list(self.q_ij[i].keys()) != want:
                raise ValueError('coupling constraints of agent %d do not share the full vehicle splines '
                                 'with exactly its fleet neighbours: the batched ADMM does not cover it' % i)

    def _build_consensus_projector(self):
        """Coupling constraints of one agent's z-update (formation.py:47-65 seen
        through admm.py:313-354): z = [z_i, z_ij (per neighbour)];
        centre_i - centre_j = 0 for every couple the agent belongs to and
        centre_i^(d)(1) = 0, d = 1..degree."""
        L, ns, nn = self.L, self.ns, self.n_nghb
        nz = self.nsh * (1 + nn)
        syms = pl.sym_array('zc', 'var', nz, 1)[:, 0]
        ids = [s.single_symbol() for s in syms]
        rel = pl.sym_array('relc', 'par', ns * (1 + nn), 1)[:, 0]
        rel_ids = [r.single_symbol() for r in rel]

        def centre(block, k):
            off = block * self.nsh + k * L
            return BSpline(self.basis, syms[off:off + L]) + rel[block * ns + k]

        rows = []
        for j in range(nn):                 # couples (i, nghb_j): centre_i - centre_j
            for k in range(ns):
                rows += list((centre(0, k) - centre(1 + j, k)).coeffs)
        for k in range(ns):                 # own terminal derivative constraints
            c = centre(0, k)
            for d in range(1, self.basis.degree + 1):
                rows.append(c.derivative(d)(1.)[0])
        self._rel_ids = rel_ids
        self._coupling_rows = rows
        self._z_ids = ids
        # A does not depend on the parameters (rel_pos enters b only)
        A, _ = _linear_rows(rows, ids, {r: 0.0 for r in rel_ids})
        self.A = A
        # row-normalised copy for the projector (derivative rows are ~1e3 larger)
        self._rown = 1.0 / np.linalg.norm(A, axis=1)
        An = A * self._rown[:, None]
        self.M = An.T.dot(np.linalg.inv(An.dot(An.T)))  # nz x n_con (acts on normalised b)
        self.Pz = np.eye(nz) - self.M.dot(An)
        self.nz = nz

    def _b_of(self, i):
        """Right-hand side b_i of agent i's coupling constraints (A z = b)."""
        vals = {}
        relv = [self.vehicles[i].rel_pos_c] + [self.vehicles[j].rel_pos_c for j in self.nghb[i]]
        for blk, rv in enumerate(relv):
            for k in range(self.ns):
                vals[self._rel_ids[blk * self.ns + k]] = float(rv[k])
        _, b = _linear_rows(self._coupling_rows, self._z_ids, vals)
        return b

    def _init_agent_data(self):
        N, nsh, nn = self.N, self.nsh, self.n_nghb
        f = self.father
        n, n_par = self.tb.n, self.tb.n_par
        self.X = np.zeros((N, n))
        self.P = np.zeros((N, n_par))
        self.c = np.zeros((N, self.nz))
        for i, v in enumerate(self.vehicles):
            self.X[i] = self._cold_start(v)
            self.c[i] = self.M.dot(self._rown * self._b_of(i))
        self.x_i = self.X[:, self.x_off:self.x_off + nsh].copy()
        self.z_i = self.x_i.copy()                       # admm.py:356-366 init_var_admm
        self.l_i = np.zeros((N, nsh))
        self.x_j = self.x_i[self.nghb]                   # (N, nn, nsh)
        self.z_ij = np.zeros((N, nn, nsh))
        self.l_ij = np.zeros((N, nn, nsh))
        self.z_ji = self.x_i[:, None, :].repeat(nn, 1)   # z_ji initialised with own x
        self.l_ji = np.zeros((N, nn, nsh))
        self.state = np.array([v.prediction['state'] for v in self.vehicles], dtype=float)
        self.inp = np.array([v.prediction['input'] for v in self.vehicles], dtype=float)
        self.poseT = np.array([v.poseT for v in self.vehicles], dtype=float)
        self.relp = np.array([v.rel_pos_c for v in self.vehicles], dtype=float)

    def _cold_start(self, vehicle):
        f = self.father
        x = f._var_struct(0.)
        x[self.veh_label, 'splines_seg0'] = vehicle.get_init_spline_value()[0]
        return x.cat

    # ------------------------------------------------------------------
    def pack_parameters(self, t):
        P, off = self.P, self.par_off
        base = self.father.set_parameters(0.).cat
        P[:] = base[None]
        v, a = self.veh_label, self.upd_label
        N, nsh, nn = self.N, self.nsh, self.n_nghb
        P[:, off[(v, 'rel_pos_c')]:off[(v, 'rel_pos_c')] + self.ns] = self.relp
        P[:, off[(v, 'state0')]:off[(v, 'state0')] + 2] = self.state
        P[:, off[(v, 'input0')]:off[(v, 'input0')] + 2] = self.inp
        P[:, off[(v, 'poseT')]:off[(v, 'poseT')] + 2] = self.poseT
        P[:, off[(self.p2p_label, 't')]] = np.round(t, 6) % self.knot_time
        P[:, off[(self.p2p_label, 'T')]] = self.T
        P[:, off[(a, 'z_i')]:off[(a, 'z_i')] + nsh] = self.z_i
        P[:, off[(a, 'z_ji')]:off[(a, 'z_ji')] + nsh * nn] = self.z_ji.reshape(N, -1)
        P[:, off[(a, 'l_i')]:off[(a, 'l_i')] + nsh] = self.l_i
        P[:, off[(a, 'l_ji')]:off[(a, 'l_ji')] + nsh * nn] = self.l_ji.reshape(N, -1)
        P[:, off[(a, 'rho')]] = self.options['rho']
        return P

    def first_knot_transforms(self, t):
        t0 = (np.round(t, 6) % self.knot_time) / self.T
        Tf, Tb = shiftfirstknot_T(self.basis, t0, inverse=True)
        return np.asarray(Tf, dtype=float), np.asarray(Tb, dtype=float)

    # ------------------------------------------------------------------
    # host reference of the z/l/residual update (used by the oracle tests; the
    # product path is the CUDA kernel omg_admm_zl_update via solver/b200.py)
    # ------------------------------------------------------------------
    def shift_over_knot(self):
        Ts = shiftoverknot_T(self.basis)
        L = self.L

        def tf(arr):
            shp = arr.shape
            a = arr.reshape(-1, L)
            return a.dot(Ts.T).reshape(shp)
        for key in ('x_i', 'z_i', 'l_i', 'x_j', 'z_ij', 'l_ij', 'z_ji', 'l_ji'):
            setattr(self, key, tf(getattr(self, key)))
        for blk in self.father.shifted_entries():
            _, _, off, shape, T = blk
            for c in range(shape[1]):
                seg = slice(off + c * shape[0], off + (c + 1) * shape[0])
                self.X[:, seg] = self.X[:, seg].dot(np.asarray(T).T)



class RendezVous(FormationPoint2point):
    """Vehicles that must agree on WHERE to meet: every agent solves a FreeEndPoint2point
    (its terminal conditions conT0 are variables) and the ADMM consensus runs on conT0 +
    rel_pos_c instead of on whole trajectories (reference ``omgtools/problems/
    rendezvous.py``: agents 28-35, coupling constraints 37-61).  Same batched machinery as
    FormationPoint2point; the shared block has no spline structure (block length 1, identity
    first-knot transforms, nothing to shift over a knot)."""

    def init(self, build_solver=True):
        from .point2point import FreeEndPoint2point
        veh = self.vehicles[0]
        p2p_opts = {k: v for k, v in self.options.items()
                    if k in ('horizon_time', 'solver', 'solver_options', 'verbose')}
        free_ind = sorted(self.fleet.configuration[veh].keys())
        self.p2p = FreeEndPoint2point(veh, self.environment.copy(), p2p_opts, {veh: free_ind})
        self.updater = ADMMUpdater()
        env = self.p2p.environment
        father = OptiFather([veh, self.p2p, env, self.updater] + list(env.obstacles))
        father.reset()
        self.rel_pos_c = veh.define_parameter('rel_pos_c', len(free_ind))
        self.p2p.father = father
        self.p2p.construct()
        conT = self.p2p._variables['conT0']
        self.updater.construct(veh, self.p2p, self.n_nghb, self.options['AMA'], shared=conT)
        self.father = father
        if build_solver:
            self.solver, _ = father.construct_problem(self.p2p.options)
        else:
            father.translate_symbols()
            father.construct_variables()
            father.construct_parameters()
            rows, lb, ub = father.construct_constraints()
            from ..basics.lowering import lower
            father.tables = lower(father._var_ids, father._par_ids, rows,
                                  father.construct_objective(), lb, ub, father.order_hint())
            father.init_variables()
            father.init_parameters()
            self.solver = None
        father.init_transformations(self.p2p.init_primal_transform, self.p2p.init_dual_transform)
        self.tb = father.tables
        self.basis = veh.basis
        self.L, self.ns = 1, len(free_ind)            # shared blocks of length 1
        self.nsh = len(free_ind)
        self.knot_time = self.p2p.knot_time
        self.T = self.options['horizon_time']
        self.x_off = father._var_struct.entries[(self.p2p.label, 'conT0')][0]
        self.par_off = {k: v[0] for k, v in father._par_struct.entries.items()}
        self.veh_label, self.p2p_label, self.upd_label = veh.label, self.p2p.label, self.updater.label
        self._build_consensus_projector()
        self._init_agent_data()

    def _build_consensus_projector(self):
        """conT_i + r_i - (conT_j + r_j) = 0 for every neighbour j (rendezvous.py:47-58)."""
        ns, nn = self.ns, self.n_nghb
        nz = ns * (1 + nn)
        syms = pl.sym_array('zc', 'var', nz, 1)[:, 0]
        rel = pl.sym_array('relc', 'par', nz, 1)[:, 0]
        rows = [(syms[k] + rel[k]) - (syms[(1 + j) * ns + k] + rel[(1 + j) * ns + k])
                for j in range(nn) for k in range(ns)]
        self._rel_ids = [r.single_symbol() for r in rel]
        self._coupling_rows = rows
        self._z_ids = [s.single_symbol() for s in syms]
        A, _ = _linear_rows(rows, self._z_ids, {r: 0.0 for r in self._rel_ids})
        self.A = A
        self._rown = 1.0 / np.linalg.norm(A, axis=1)
        An = A * self._rown[:, None]
        self.M = An.T.dot(np.linalg.inv(An.dot(An.T)))
        self.Pz = np.eye(nz) - self.M.dot(An)
        self.nz = nz

    def _cold_start(self, vehicle):
        x = self.father._var_struct(0.)
        x[self.veh_label, 'splines_seg0'] = vehicle.get_init_spline_value()[0]
        x[self.p2p_label, 'conT0'] = np.asarray(vehicle.poseT, dtype=float)[:self.ns]
        return x.cat

    def first_knot_transforms(self, t):
        return np.eye(1), np.eye(1)

    def shared_shift_T(self):
        return np.eye(1)
