"""Formation ADMM (BASELINE config 3): structure, consensus projector and the
multi-rank neighbour exchange (gloo, world_size 2, CPU)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from omg_tools_b200 import scenarios as sc
from omg_tools_b200.problems.admm_gpu import AgentExchange


@pytest.fixture(scope='module')
def formation():
    return sc.config3(4, build_solver=False)


def test_agent_nlp_dimensions(formation):
    tb = formation.tb
    # per-agent x-update NLP
    assert (tb.n, tb.m, tb.n_par) == (118, 578, 203)
    names = [k[1] for k in formation.father._par_struct.keys()]
    assert names[:6] == ['rel_pos_c', 'state0', 'input0', 'poseT', 'T', 't']
    assert names[6:11] == ['z_i', 'z_ji', 'l_i', 'l_ji', 'rho']
    assert formation.A.shape == (58, 78) and np.linalg.matrix_rank(formation.A) == 58
    # objective is quadratic in the shared variables: Hessian terms on the objective row
    assert np.any(tb.W.lrow == tb.m)


def test_projector_equals_reference_kkt_solve(formation):
    """z = P v + c  ==  the reference's Schur-complement z-update (admm.py:149-155)."""
    p = formation
    rng = np.random.default_rng(0)
    rho = 1.7
    for i in range(p.N):
        x, l = rng.standard_normal(p.nz), rng.standard_normal(p.nz)
        b = p._b_of(i)
        f = -(l + rho * x)
        G = -(1. / rho) * p.A.dot(p.A.T)
        h = b + (1. / rho) * p.A.dot(f)
        mu = np.linalg.solve(G, h)
        z_ref = -(1. / rho) * (p.A.T.dot(mu) + f)
        z = p.Pz.dot(x + l / rho) + p.c[i]
        assert np.abs(z - z_ref).max() < 1e-9
        assert np.abs(p.A.dot(z) - b).max() < 1e-8


def test_neighbour_tables(formation):
    p = formation
    for i in range(p.N):
        for k, j in enumerate(p.nghb[i]):
            assert p.nghb[j][p.back[i, k]] == i


def _worker(rank, world, port, N, nsh, nn, nghb, back, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    x = torch.tensor(rng.standard_normal((N, nsh)))
    z = torch.tensor(rng.standard_normal((N, nn, nsh)))
    l = torch.tensor(rng.standard_normal((N, nn, nsh)))
    ex = AgentExchange(N, nghb, back, rank, world)
    xj = ex.gather_x(x[ex.lo:ex.hi].clone())
    zji, lji = ex.gather_zl(z[ex.lo:ex.hi].clone(), l[ex.lo:ex.hi].clone())
    tot = ex.allreduce_sum(torch.tensor([float(rank + 1), 2., 3.], dtype=torch.float64))
    torch.save({'xj': xj, 'zji': zji, 'lji': lji, 'tot': tot, 'lo': ex.lo, 'hi': ex.hi},
               os.path.join(out, 'r%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_two_ranks_gloo(tmp_path):
    N, nsh, nn = 8, 26, 2
    nghb = np.array([[(i + 1) % N, (i - 1) % N] for i in range(N)])
    back = np.array([[list(nghb[j]).index(i) for j in nghb[i]] for i in range(N)])
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, N, nsh, nn, nghb, back, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((N, nsh))
    z = rng.standard_normal((N, nn, nsh))
    l = rng.standard_normal((N, nn, nsh))
    for rank in range(2):
        d = torch.load(os.path.join(str(tmp_path), 'r%d.pt' % rank))
        lo, hi = d['lo'], d['hi']
        assert (lo, hi) == (rank * 4, rank * 4 + 4)
        for i in range(lo, hi):
            for k, j in enumerate(nghb[i]):
                assert np.array_equal(d['xj'][i - lo, k].numpy(), x[j])
                assert np.array_equal(d['zji'][i - lo, k].numpy(), z[j, back[i, k]])
                assert np.array_equal(d['lji'][i - lo, k].numpy(), l[j, back[i, k]])
        assert d['tot'].tolist() == [3., 4., 6.]


def test_admm_oracle_reduces_formation_error(formation):
    from oracle.admm_ref import ADMMOracle
    orc = ADMMOracle(sc.config3(4, build_solver=False))
    spread = []
    for _ in range(8):
        res = orc.dual_update(0.)
        cen = orc.x_i.reshape(4, 2, 13) + orc.p.relp[:, :, None]
        spread.append(np.abs(cen - cen.mean(0)).max())
        assert np.all(orc.status == 0)
    assert spread[-1] < 0.2 * spread[0]


def test_admm_oracle_nesterov_acceleration(formation):
    """Fast ADMM (reference admm.py:510-554): the extrapolated iteration differs
    from plain ADMM from the second iteration on and still drives the agents to
    consensus; with nesterov_reset the step is undone when the combined residual
    grows."""
    from oracle.admm_ref import ADMMOracle
    plain = ADMMOracle(sc.config3(4, build_solver=False))
    fast = ADMMOracle(sc.config3(4, {'nesterov_acceleration': True}, build_solver=False))
    reset = ADMMOracle(sc.config3(4, {'nesterov_acceleration': True, 'nesterov_reset': True},
                                  build_solver=False))
    hist = {k: [] for k in ('plain', 'fast', 'reset')}
    for k in range(8):
        for name, orc in (('plain', plain), ('fast', fast), ('reset', reset)):
            hist[name].append(orc.dual_update(0.))
            assert np.all(orc.status == 0)
        if k == 0:      # alpha_0 = 1: the first extrapolation weight is zero
            assert np.allclose(fast.z_i, plain.z_i) and np.allclose(fast.l_i, plain.l_i)
    assert not np.allclose(fast.z_i, plain.z_i)
    assert fast.alpha > 3.0 and reset.alpha <= fast.alpha
    for name in ('fast', 'reset'):
        pr = np.array([h[0] for h in hist[name]])
        assert pr[-1] < 0.3 * pr[0]          # primal residual (consensus error) shrinks


@pytest.mark.parametrize('interconnection', ['circular', 'line'])
def test_admm_converges_to_the_centralised_optimum(interconnection):
    """Independent check of the whole ADMM machinery (x-update NLP with the augmented
    Lagrangian, consensus projector, multiplier update, neighbour exchange): its fixed
    point must be the optimum of the COUPLED problem -- all four vehicles in one NLP with the
    formation constraints x_i + r_i = x_j + r_j as equality rows (what the reference's
    FormationPoint2pointCentral states, formation_central.py:36-78, here with one terminal
    slack per vehicle so that the objective is the sum of the agents' objectives).  The
    coupled NLP (n = 472, m = 2390) is solved by the C oracle; 100 ADMM iterations bring every
    agent's spline coefficients to within 2 cm of it.  'line': an open chain, the end vehicles
    have ONE neighbour (unequal neighbour counts: the missing slot holds a copy of the agent
    itself, problems/admm.py) -- the same fixed point, a little slower."""
    from oracle import ipm_c
    from oracle.admm_ref import ADMMOracle
    if not ipm_c.available():
        pytest.skip('C oracle not built')
    from omg_tools_b200 import (Holonomic, Fleet, Environment, Obstacle, Rectangle, Square,
                                RegularPolyhedron)
    from omg_tools_b200.problems.point2point import FixedTPoint2point
    from omg_tools_b200.basics.optilayer import inf
    from omg_tools_b200.basics.spline_extra import definite_integral
    from omg_tools_b200.basics.lowering import lower

    class Coupled(FixedTPoint2point):
        def define_terminal_constraints(self):
            objective = 0.
            for v, vehicle in enumerate(self.vehicles):
                term_con, term_con_der = vehicle.get_terminal_constraints(vehicle.splines[0])
                for k, (spline, condition) in enumerate(term_con):
                    g = self.define_spline_variable('g%d_%d' % (v, k), 1, basis=spline.basis)[0]
                    objective += definite_integral(g, self.t0, 1.)
                    self.define_constraint(spline - condition - g, -inf, 0.)
                    self.define_constraint(-spline + condition - g, -inf, 0.)
                for spline, condition in term_con_der:
                    self.define_constraint(spline(1.) - condition, 0., 0.)
            self.define_objective(objective)

        def construct(self):
            FixedTPoint2point.construct(self)
            for a, b in zip(self.vehicles[:-1], self.vehicles[1:]):     # a chain: no redundant rows
                for d in range(2):
                    self.define_constraint((a.splines[0][d] + a.rel_pos_c[d]) -
                                           (b.splines[0][d] + b.rel_pos_c[d]), 0., 0.)

    N = 4
    vehicles = [Holonomic() for _ in range(N)]
    fleet = Fleet(vehicles)
    configuration = RegularPolyhedron(0.2, N, np.pi / 4.).vertices.T
    fleet.set_configuration(configuration.tolist())
    fleet.set_initial_conditions((np.array([-1.5, -1.5]) + configuration).tolist())
    fleet.set_terminal_conditions((np.array([2., 2.]) + configuration).tolist())
    environment = Environment(room={'shape': Square(5.)})
    rectangle = Rectangle(width=3., height=0.2)
    environment.add_obstacle(Obstacle({'position': [-2.1, -0.5]}, shape=rectangle))
    environment.add_obstacle(Obstacle({'position': [1.7, -0.5]}, shape=rectangle))
    pr = Coupled(fleet, environment, {'verbose': 0, 'horizon_time': 10})
    f = pr.father
    f.reset()
    pr.construct()
    f.translate_symbols()
    f.construct_variables()
    f.construct_parameters()
    rows, lb, ub = f.construct_constraints()
    tb = f.tables = lower(f._var_ids, f._par_ids, rows, f.construct_objective(), lb, ub, f.order_hint())
    f.init_variables()
    f.init_parameters()
    f.init_transformations(pr.init_primal_transform, pr.init_dual_transform)
    pr.reinitialize()
    assert (tb.n, tb.m) == (4 * 118, 2390)
    r = ipm_c.solve_batch_full(tb, f.get_variables().cat[None], f.set_parameters(0.).cat[None], threads=1,
                               options={'tol': 1e-6, 'compl_inf_tol': 1e-7, 'constr_viol_tol': 1e-7})
    assert r['status'][0] == 0
    ent = f._var_struct.entries
    central = np.array([r['x'][0][ent[(v.label, 'splines_seg0')][0]:][:26] for v in vehicles])
    prd = sc.config3(N, build_solver=False, interconnection=interconnection)
    if interconnection == 'line':
        assert prd.n_nghb == 2 and (~prd.real_nghb).sum() == 2 and prd.nghb[0, 1] == 0 and prd.nghb[N - 1, 0] == N - 1
    orc = ADMMOracle(prd)
    err = []
    for k in range(100 if interconnection == 'circular' else 150):
        p_res, d_res, c_res = orc.dual_update(0.)
        err.append(np.abs(orc.x_i - central).max())
    print(interconnection, err[0], err[-1], min(err[-10:]), p_res)
    assert err[0] > 0.4 and err[-1] < 0.02 and min(err[-10:]) < 0.01
    assert p_res < 0.02


def test_rendezvous_reaches_consensus():
    """RendezVous (reference rendezvous.py): agents solving FreeEndPoint2point problems agree
    by ADMM on their terminal positions conT0 + rel_pos_c; the shared block has no spline
    structure (block length 1, identity transforms).  Sixteen iterations bring the proposed
    meeting centres of four vehicles from 0.5 m apart into the millimetre range -- the floor the
    x-updates' own tolerance (tol = 1e-3, problem.py:57) sets: there the spread wanders between
    1e-5 and a few 1e-3 from one iteration to the next, so the bound is on the last iterations
    together -- every x-update converged, and the agreed point satisfies the coupling constraints."""
    from oracle import ipm_c
    from oracle.admm_ref import ADMMOracle
    if not ipm_c.available():
        pytest.skip('C oracle not built')
    pr = sc.config_rendezvous(4, build_solver=False)
    assert (pr.tb.n, pr.nsh, pr.L, pr.A.shape) == (87, 2, 1, (4, 6))
    names = [k[1] for k in pr.father._var_struct.keys()]
    assert 'conT0' in names
    orc = ADMMOracle(pr)
    spread = []
    for _ in range(16):
        p_res, d_res, c_res = orc.dual_update(0.)
        assert np.all(orc.status == 0)
        centre = orc.x_i + pr.relp
        spread.append(np.abs(centre - centre.mean(0)).max())
    assert spread[0] > 0.3 and max(spread[-6:]) < 5e-3 and min(spread[-6:]) < 1e-3 and p_res < 5e-3
    for i in range(pr.N):
        z = np.r_[orc.z_i[i], orc.z_ij[i].reshape(-1)]
        assert np.abs(pr.A.dot(z) - pr._b_of(i)).max() < 1e-9
    # every vehicle's trajectory ends at its agreed terminal position
    L = len(pr.basis)
    ends = orc.X[:, [L - 1, 2 * L - 1]]
    assert np.abs(ends - orc.x_i).max() < 1e-2


def test_admm_ama_option(formation):
    """Option 'AMA' (alternating minimisation, reference admm.py:97-104): the x-update
    drops the quadratic penalty -- the agent NLP's objective becomes linear in x (no
    Hessian terms on the objective row) -- and the fast variant extrapolates the
    multipliers only (admm.py:527-541).  The holonomic formation objective is not strongly
    convex, so no convergence claim is made here (the reference uses AMA for the quadrotor
    comparison only); the iteration must run and keep the consensus constraint A z = b."""
    from oracle.admm_ref import ADMMOracle
    pr = sc.config3(4, {'AMA': True, 'nesterov_acceleration': True}, build_solver=False)
    tb = pr.tb
    assert (tb.n, tb.m, tb.n_par) == (118, 578, 203)
    assert not np.any(tb.W.lrow == tb.m) and np.any(formation.tb.W.lrow == formation.tb.m)
    orc = ADMMOracle(pr)
    for _ in range(3):
        res = orc.dual_update(0.)
        assert np.all(np.isfinite(res))
    for i in range(pr.N):
        z = np.r_[orc.z_i[i], orc.z_ij[i].reshape(-1)]
        assert np.abs(pr.A.dot(z) - pr._b_of(i)).max() < 1e-8


def test_fleet_configuration_equals_the_references():
    """Fleet.set_configuration / get_neighbors (vehicles/fleet.py) of the reference,
    run from /root/reference (tests/golden/make_model_golden.py): relative positions
    to the formation centre and the circular neighbour lists."""
    import os
    from omg_tools_b200 import Holonomic, Fleet, RegularPolyhedron
    M = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'model_golden.npz'))
    for n_agents in (4, 6):
        conf = RegularPolyhedron(0.2, n_agents, np.pi / 4.).vertices.T
        fleet = Fleet([Holonomic() for _ in range(n_agents)])
        fleet.set_configuration(conf.tolist())
        rel = np.array([v.rel_pos_c for v in fleet.vehicles], float)
        nghb = np.array([[fleet.vehicles.index(w) for w in fleet.get_neighbors(v)]
                         for v in fleet.vehicles])
        assert np.abs(rel - M['fleet%d_rel_pos_c' % n_agents]).max() < 1e-15
        assert np.array_equal(nghb, M['fleet%d_nghb' % n_agents])


@pytest.mark.parametrize('interconnection', ['circular', 'line'])
def test_interprete_constraints_derives_the_shared_sets(interconnection):
    """problems/distributed.py vs the reference's DistributedProblem.interprete_constraints
    (distributedproblem.py:105-169): from the formation constraints centre_i - centre_j = 0 of
    neighbouring vehicles (formation.py:47-65) the shared sets come out as ALL spline coefficients
    of every vehicle, copies of exactly the neighbours' coefficients -- also for a fleet whose
    vehicles have different numbers of neighbours and for a constraint that couples only a
    subset of the coefficients of a vehicle of another type."""
    from omg_tools_b200 import Holonomic, Fleet
    from omg_tools_b200.basics.spline import BSpline
    from omg_tools_b200.problems.distributed import variable_owners, interprete_constraints
    N = 4
    vehicles = [Holonomic() for _ in range(N)]
    fleet = Fleet(vehicles, interconnection=interconnection)
    fleet.set_configuration([[0.2, 0.], [0., 0.2], [-0.2, 0.], [0., -0.2]])
    splines = []
    for veh in vehicles:
        veh.reset() if hasattr(veh, 'reset') else None
        coeffs = veh.define_variable('splines_seg0', len(veh.basis), veh.n_spl)
        splines.append([BSpline(veh.basis, coeffs[:, k]) for k in range(veh.n_spl)])
    cons = []
    for i, veh in enumerate(vehicles):
        for other in fleet.get_neighbors(veh):
            j = vehicles.index(other)
            for k in range(2):
                cons.append(((splines[i][k] + veh.rel_pos_c[k]) - (splines[j][k] + other.rel_pos_c[k])).coeffs)
    owners = variable_owners([[v] for v in vehicles])
    q_i, q_ij, q_ji = interprete_constraints(owners, cons)
    L = len(vehicles[0].basis)
    for i, veh in enumerate(vehicles):
        assert list(q_i[i].keys()) == [veh.label] and q_i[i][veh.label]['splines_seg0'] == list(range(2 * L))
        nghb = sorted(vehicles.index(w) for w in fleet.get_neighbors(veh))
        assert list(q_ij[i].keys()) == nghb and list(q_ji[i].keys()) == nghb
        for j in nghb:
            assert q_ij[i][j][vehicles[j].label]['splines_seg0'] == list(range(2 * L))
            assert q_ji[i][j] is q_ij[j][i]
    if interconnection == 'line':
        assert [len(q) for q in q_ij] == [1, 2, 2, 1]
    # a partial coupling: only the last coefficient of x of vehicle 0 with a scalar of vehicle 3
    s = np.asarray(vehicles[3].define_variable('meet', 1), dtype=object).reshape(-1)
    q_i, q_ij, _ = interprete_constraints(variable_owners([[v] for v in vehicles]),
                                          [splines[0][0].coeffs[-1] - s[0]])
    assert q_i[0][vehicles[0].label]['splines_seg0'] == [L - 1] and q_i[3][vehicles[3].label]['meet'] == [0]
    assert q_ij[0][3][vehicles[3].label]['meet'] == [0] and q_ij[3][0][vehicles[0].label]['splines_seg0'] == [L - 1]


def test_dual_decomposition_subproblem_and_ascent():
    """Dual decomposition (reference dualdecomposition.py / formation_dualdec.py): every agent's
    NLP returns its own trajectory and copies of its neighbours' that satisfy the formation rows
    exactly; with the multipliers at zero the agents ignore each other (residual = the mismatch of
    the uncoordinated plans); the dual ascent with a small step then drives the residual down --
    a sub-gradient method on a piecewise-linear dual, so plateaus and jumps, not a monotone
    sequence (the reference runs it with rho = 0.003, examples/compare_distributed_optimization_
    quadrotors.py:95)."""
    from oracle import ipm_c
    from oracle.admm_ref import DDOracle
    if not ipm_c.available():
        pytest.skip('C oracle not built')
    pr = sc.config_formation_dd(4, build_solver=False, options={'rho': 0.02})
    assert (pr.tb.n, pr.tb.kkt_n_eq) == (118 + 52, 10 + 52)
    orc = DDOracle(pr)
    res = []
    for it in range(16):
        res.append(orc.dual_update(0.))
        assert np.all(orc.status == 0)
        # formation rows inside every agent's NLP: centre_i(x_i) + r_i = centre_j(z_ij) + r_j
        ci = orc.x_i.reshape(pr.N, 1, pr.ns, pr.L) + pr.relp[:, None, :, None]
        cj = orc.z_ij.reshape(pr.N, pr.n_nghb, pr.ns, pr.L) + pr.relp[pr.nghb][:, :, :, None]
        assert np.abs(ci - cj).max() < 1e-6
    assert res[0] > 1.0 and min(res[8:]) < 0.35 * res[0]
    # multipliers a pair holds for each other are exchanged consistently
    assert np.array_equal(orc.l_ji, orc.l_ij[pr.nghb, pr.back])
