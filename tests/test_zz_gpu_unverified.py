"""GPU parity tests of kernel paths written after the round's GPU budget was spent: they
compile for sm_100a and pass in the CPU emulation of the kernel source
(tests/test_kernel_emulation.py, tools/cpu_emu), but have NOT run on a GPU yet.  They are
non-strict xfail so that a first GPU run reports XPASS / XFAIL without masking the verified
suite (this file sorts last); drop the marker once XPASS."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason='kernel path not yet run on a GPU '
                                '(cross-Hessian slots of the XL kernel, ABI v6)')]

from omg_tools_b200 import scenarios as sc
from oracle import ipm_c

NORTH_STAR_TOL = 1e-4


def test_dubins_default_formulation_matches_oracle():
    """Dubins without substitution (dubins.py:63, 235-251): rows affine in the shared
    intermediates with x-dependent coefficients -> cross-Hessian slots X and the gather
    X^T C + C^T X in the XL kernel (csrc/omg_b200.cu) vs oracle/ipm.c on 8 jittered
    instances: same statuses, iteration counts within 2, flat-output splines within the
    north-star tolerance."""
    pr = sc.config_dubins_plain()
    tb = pr.father.tables
    assert tb.nnz_wx > 0
    X0, P = sc.instance_data(pr, 8, jitter=0.1, seed=1)
    res = pr.problem.solve_batch(X0, P)
    ref = ipm_c.solve_batch_full(tb, X0, P, threads=8)
    assert np.array_equal(res['status'], ref['status'])
    ok = ref['status'] == 0
    assert ok.sum() >= 7
    assert np.abs(res['iters'] - ref['iters'])[ok].max() <= 2
    err = np.abs(res['x'] - ref['x'])[ok][:, :26].max(axis=1)       # v~ and tan(theta/2) splines
    assert np.median(err) < NORTH_STAR_TOL
    assert np.abs(res['f'] - ref['f'])[ok].max() < 1e-3


def test_dubins_default_formulation_problem_solve_dropin():
    """The reference-facing call Problem.solve() on the same problem."""
    pr = sc.config_dubins_plain()
    pr.initialize(0.)
    pr.solve(0., 0.5)
    assert pr.problem.stats()['return_status'] == 'Solve_Succeeded'


def test_holonomic_orient_matches_oracle():
    """HolonomicOrient (m = 3035 rows, 232 shared heading products): the XL kernel with the
    cross-Hessian gather at a row count no verified test reaches."""
    pr = sc.config_holonomic_orient()
    tb = pr.father.tables
    X0, P = sc.instance_data(pr, 4, jitter=0.05, seed=2)
    res = pr.problem.solve_batch(X0, P)
    ref = ipm_c.solve_batch_full(tb, X0, P, threads=4)
    assert np.array_equal(res['status'], ref['status'])
    ok = ref['status'] == 0
    assert ok.sum() >= 3
    err = np.abs(res['x'] - ref['x'])[ok][:, :39].max(axis=1)
    assert np.median(err) < 1e-3
    assert np.abs(res['f'] - ref['f'])[ok].max() < 1e-3


def test_bicycle_mid_mid_hessian_matches_oracle():
    """Bicycle (vehicles/bicycle.py): rows with products of two shared product splines ->
    the C^T M C gather of the XL kernel, from a rolling initial guess."""
    pr = sc.config_bicycle()
    tb = pr.father.tables
    X0, P = sc.instance_data(pr, 4, jitter=0.02, seed=3)
    X0[:, :7] = 0.3
    res = pr.problem.solve_batch(X0, P)
    ref = ipm_c.solve_batch_full(tb, X0, P, threads=4)
    both = (res['status'] == 0) & (ref['status'] == 0)
    assert both.sum() >= 2
    assert res['status'][0] == 0 and res['iters'][0] == ref['iters'][0]
    assert np.abs(res['x'] - ref['x'])[0].max() < 1e-4


def test_simple_quadrotor3d_matches_oracle():
    """SimpleQuadrotor3D (standard kernel, 1 block/SM layout: 226 KB of shared memory)."""
    pr = sc.config_quadrotor3d_simple()
    tb = pr.father.tables
    X0, P = sc.instance_data(pr, 4, jitter=0.05, seed=1)
    res = pr.problem.solve_batch(X0, P)
    ref = ipm_c.solve_batch_full(tb, X0, P, threads=4)
    assert np.array_equal(res['status'], ref['status']) and res['status'][0] == 0
    ok = ref['status'] == 0
    assert np.median(np.abs(res['x'] - ref['x'])[ok][:, :42].max(axis=1)) < NORTH_STAR_TOL


def test_rendezvous_admm_matches_oracle():
    """RendezVous on the GPU runner (shared blocks of length 1 in the consensus kernel) vs
    the sequential ADMM oracle, iteration by iteration."""
    from omg_tools_b200.problems.admm_gpu import FormationADMMRunner
    from oracle.admm_ref import ADMMOracle
    run = FormationADMMRunner(sc.config_rendezvous(4))
    orc = ADMMOracle(sc.config_rendezvous(4, build_solver=False))
    for it in range(8):
        rg = run.dual_update(0.)
        ro = orc.dual_update(0.)
        st, _ = run.status()
        assert np.all(st == 0) and np.all(orc.status == 0)
        assert np.abs(run.x_i.cpu().numpy() - orc.x_i).max() < NORTH_STAR_TOL, it
        assert np.abs(run.z_i.cpu().numpy() - orc.z_i).max() < NORTH_STAR_TOL
        assert np.abs(run.l_i.cpu().numpy() - orc.l_i).max() < 10 * NORTH_STAR_TOL
        assert abs(rg[0] - ro[0]) < 1e-3 * max(1., ro[0])


def test_trailer_matches_oracle():
    """Trailer + Dubins lead vehicle (2.1 M Jacobian terms, T x intermediate cross terms)."""
    pr = sc.config_trailer(init_v_til=0.3)
    tb, f = pr.father.tables, pr.father
    X0, P = f.get_variables().cat[None], f.set_parameters(0.).cat[None]
    res = pr.problem.solve_batch(X0, P)
    ref = ipm_c.solve_batch_full(tb, X0, P, threads=1)
    assert res['status'][0] == 0 == ref['status'][0]
    assert abs(int(res['iters'][0]) - int(ref['iters'][0])) <= 2
    assert np.abs(res['x'] - ref['x'])[:, :36].max() < 1e-3


def test_feasibility_kernel_matches_oracle():
    """omg_feas_kernel (written after the GPU budget was spent; emulation-verified in
    tests/test_kernel_emulation.py) vs oracle_feas_batch: standard tables (config 5) and
    tables with intermediates (Dubins, free end time), 16 jittered cold starts each."""
    for name, seed in (('config5', 5), ('config_dubins_freeT', 3)):
        pr = getattr(sc, name)()
        tb = pr.father.tables
        X0, P = sc.instance_data(pr, 16, jitter=0.2, seed=seed)
        xg, vg, kg = pr.problem.feasibility_batch(X0, P)
        xc, vc, kc = ipm_c.feas_batch(tb, X0, P)
        assert np.array_equal(kg, kc), name
        # FMA contraction on the GPU: agreement to rounding amplified by 30 LM steps
        assert np.abs(vg - vc).max() < 1e-6 * max(1., np.abs(vc).max()), name
        assert np.abs(xg - xc).max() < 1e-5, name


def test_dubins_example_as_written_converges_through_the_feasibility_phase():
    """examples/p2p_dubins.py from the reference's zero-speed guess: Restoration_Failed
    after a few iterations, feasibility phase, second solve -> end time 7.46 s, through
    B200Solver.solve_batch and through the reference-facing Problem.solve()."""
    pr = sc.config_dubins_freeT()
    tb, f = pr.father.tables, pr.father
    X0, P = f.get_variables().cat[None], f.set_parameters(0.).cat[None]
    res = pr.problem.solve_batch(X0, P)
    ref = ipm_c.solve_batch_full(tb, X0, P, threads=1)
    assert res['status'][0] == 0 == ref['status'][0]
    assert abs(int(res['iters'][0]) - int(ref['iters'][0])) <= 2
    assert abs(res['f'][0] - ref['f'][0]) < 1e-4 and 7.0 < res['f'][0] < 8.0
    pr.initialize(0.)
    pr.solve(0., 0.5)
    assert pr.problem.stats()['return_status'] == 'Solve_Succeeded'
