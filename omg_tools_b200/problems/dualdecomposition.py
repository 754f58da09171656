"""Dual decomposition for formation point-to-point problems, batched over the agents.

Reference: ``omgtools/problems/dualdecomposition.py`` (DDUpdater: the xz-update NLP 58-147, the
multiplier update 149-171, communicate 225-230, knot shift 232-246, residual 248-258;
DDProblem.dual_update 279-314) and ``formation_dualdec.py`` (coupling constraints 32-65).

Every agent i keeps its own trajectory x_i AND copies z_ij of its neighbours' shared variables
in ONE NLP:

    min  f_i(x_i, ...)  +  sum_j  l_ji^T x_i  -  sum_j  l_ij^T z_ij
    s.t. the agent's own constraints,
         centre(x_i) + r_i - (centre(z_ij) + r_j) = 0     for every neighbour j

(x, z, l entering the objective through the future piece of their splines, ``shift_knot1_fwd``),
followed by the dual ascent  l_ij += rho (x_j - z_ij)  after the neighbours' x_j have been
communicated.  There is no z-update and no consensus projector: pure dual ascent.

As with the ADMM (problems/admm.py) all agents share ONE NLP structure, so the xz-update of all
local agents is one ``omg_solve_batch`` call whose instances differ in their parameters
(l_ij, l_ji, the neighbours' rel_pos_c); the multiplier update and the residual are elementwise
device operations; the two exchanges (x_j, then l_ji) go through the same NCCL entry point of the
C ABI as the ADMM's (``omg_admm_exchange_x``).  Equal neighbour counts are required (a padded
self-neighbour would leave its copy unconstrained and the KKT matrix singular).
"""
from __future__ import print_function

import numpy as np

from ..basics.optilayer import OptiChild, OptiFather
from ..basics.poly import Poly
from ..basics.spline import BSpline
from ..basics.spline_extra import shift_knot1_fwd
from .admm import FormationPoint2point
from .point2point import Point2point


class DDUpdater(OptiChild):
    """The copies z_ij, the multipliers and the coupling rows of one agent's xz-update
    (reference dualdecomposition.py:58-147)."""

    def __init__(self):
        OptiChild.__init__(self, 'dd')

    def construct(self, vehicle, n_nghb, rel_own):
        L, ns = len(vehicle.basis), vehicle.n_spl
        nsh = L * ns
        basis = vehicle.basis
        z_ij = np.asarray(self.define_variable('z_ij', nsh * n_nghb))
        l_ij = np.asarray(self.define_parameter('l_ij', nsh * n_nghb))
        l_ji = np.asarray(self.define_parameter('l_ji', nsh * n_nghb))
        rel_j = np.asarray(self.define_parameter('rel_pos_nghb', ns * n_nghb))
        t0 = self.define_symbol('t') / self.define_symbol('T')
        x_i = vehicle._variables['splines_seg0'].reshape(-1, order='F')     # spline 0's coefficients first

        def fwd(vec):            # future piece of every spline of a block (dualdecomposition.py:93-98)
            return np.concatenate([shift_knot1_fwd(vec[k * L:(k + 1) * L], basis, t0) for k in range(ns)])

        obj = Poly()
        xf = fwd(x_i)
        for j in range(n_nghb):
            blk = slice(j * nsh, (j + 1) * nsh)
            lji, lij, zf = fwd(l_ji[blk]), fwd(l_ij[blk]), fwd(z_ij[blk])
            for k in range(nsh):
                obj = obj + lji[k] * xf[k] - lij[k] * zf[k]
        self.define_objective(obj)
        # coupling rows (formation_dualdec.py:45-64): the neighbour's side through the copy z_ij
        rel_own = np.asarray(rel_own).reshape(-1)
        for j in range(n_nghb):
            for k in range(ns):
                ci = BSpline(basis, x_i[k * L:(k + 1) * L]) + rel_own[k]
                cj = BSpline(basis, z_ij[j * nsh + k * L:j * nsh + (k + 1) * L]) + rel_j[j * ns + k]
                self.define_constraint(ci - cj, 0., 0.)
        # (formation_dualdec.py:58-64 also imposes centre_i^(d)(1) = 0, d = 1 .. degree, on the agent's
        #  own splines.  With rel_pos_c constant these are EXACTLY the terminal-derivative rows
        #  Point2point already defines on the same splines (point2point.py:160-175): the reference
        #  hands IPOPT the duplicated equality rows and lets its delta_c regularisation cope; here
        #  the duplicates are left out -- same feasible set, full-rank equality Jacobian.)


class FormationPoint2pointDualDecomposition(FormationPoint2point):
    """Fleet of identical vehicles keeping a formation, solved by dual decomposition
    (reference formation_dualdec.py:25-66)."""

    def __init__(self, fleet, environment, options=None, rank=0, world=1, group=None):
        FormationPoint2point.__init__(self, fleet, environment, options, rank, world, group)
        if not self.real_nghb.all():
            raise ValueError('dual decomposition needs the same number of neighbours for every agent')
        self.residuals = {'primal': []}

    def init(self, build_solver=True):
        veh = self.vehicles[0]
        p2p_opts = {k: v for k, v in self.options.items()
                    if k in ('horizon_time', 'solver', 'solver_options', 'verbose',
                             'hard_term_con', 'no_term_con_der')}
        self.p2p = Point2point(veh, self.environment.copy(), p2p_opts, freeT=False)
        self.updater = DDUpdater()
        env = self.p2p.environment
        father = OptiFather([veh, self.p2p, env, self.updater] + list(env.obstacles))
        father.reset()
        self.rel_pos_c = veh.define_parameter('rel_pos_c', veh.n_dim)
        self.p2p.father = father
        self.p2p.construct()
        self.updater.construct(veh, self.n_nghb, self.rel_pos_c)
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
        self.L, self.ns = len(veh.basis), veh.n_spl
        self.nsh = self.L * self.ns
        self.knot_time = self.p2p.knot_time
        self.T = self.options['horizon_time']
        self.x_off = father._var_struct.entries[(veh.label, 'splines_seg0')][0]
        self.z_off = father._var_struct.entries[(self.updater.label, 'z_ij')][0]
        self.par_off = {k: v[0] for k, v in father._par_struct.entries.items()}
        self.veh_label, self.p2p_label, self.upd_label = veh.label, self.p2p.label, self.updater.label
        self._init_agent_data()

    def _init_agent_data(self):
        N, nsh, nn = self.N, self.nsh, self.n_nghb
        n, n_par = self.tb.n, self.tb.n_par
        self.X = np.zeros((N, n))
        self.P = np.zeros((N, n_par))
        for i, v in enumerate(self.vehicles):
            self.X[i] = self._cold_start(v)
        self.x_i = self.X[:, self.x_off:self.x_off + nsh].copy()
        self.x_j = self.x_i[self.nghb]                    # (N, nn, nsh)
        self.z_ij = self.x_j.copy()                       # dualdecomposition.py:62-70: copies start at the neighbours' guesses
        self.X[:, self.z_off:self.z_off + nsh * nn] = self.z_ij.reshape(N, -1)
        self.l_ij = np.zeros((N, nn, nsh))
        self.l_ji = np.zeros((N, nn, nsh))
        self.state = np.array([v.prediction['state'] for v in self.vehicles], dtype=float)
        self.inp = np.array([v.prediction['input'] for v in self.vehicles], dtype=float)
        self.poseT = np.array([v.poseT for v in self.vehicles], dtype=float)
        self.relp = np.array([v.rel_pos_c for v in self.vehicles], dtype=float)

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
        P[:, off[(a, 'l_ij')]:off[(a, 'l_ij')] + nsh * nn] = self.l_ij.reshape(N, -1)
        P[:, off[(a, 'l_ji')]:off[(a, 'l_ji')] + nsh * nn] = self.l_ji.reshape(N, -1)
        P[:, off[(a, 'rel_pos_nghb')]:off[(a, 'rel_pos_nghb')] + self.ns * nn] = self.relp[self.nghb].reshape(N, -1)
        return P
