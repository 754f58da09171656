"""Generate tests/golden/loop_golden.npz: the REFERENCE's receding-horizon host loop.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_loop_golden.py

The reference's ``Problem.solve()`` loop -- ``Deployer.update`` (deployer.py:43-79:
predict, solve, store) and ``Simulator.update`` (simulator.py:93-99: simulate) with
``Problem.solve`` (problem.py:103-136: init_step, get_variables, set_parameters,
update_bounds, the solver call, set_variables), ``OptiFather`` packing/unpacking and the
knot-crossing shift -- is run straight from /root/reference.  On top of the numeric
``casadi.MX`` stand-in of make_model_golden.py this needs ``casadi.tools.struct`` (a
labelled flat vector; stand-in below) and the solver object, which the reference accepts
prebuilt (``OptiFather.construct_problem(options, name, problem)``, optilayer.py:180-195):
it is THIS repository's solver stack, here the CPU oracle on the lowered tables (whose
rows equal the reference's, model_golden.npz).  Stored per MPC step: the vectors the
reference hands to the solver (x0, p, lbg, ubg) and what it unpacks (x).  tests/
test_model.py runs this framework's own loop with the same solver and must feed it
the same vectors -- i.e. the reference's ``Problem.solve()`` host path and this
framework's are interchangeable around the solver call.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_model_golden as mg                          # noqa: E402

OUT = os.path.join(HERE, 'loop_golden.npz')


# --------------------------------------------------------------------------
# stand-in for casadi.tools.struct: labelled flat vector, column-major blocks
# --------------------------------------------------------------------------
class Entry(object):
    def __init__(self, name, shape=None, struct=None, expr=None):
        self.name, self.sub = name, struct
        if struct is not None:
            self.size, self.shape = struct.size, (struct.size, 1)
        elif expr is not None:
            n = expr.a.size if isinstance(expr, mg.MX) else 1
            self.size, self.shape = n, (n, 1)
        else:
            shape = (shape, 1) if isinstance(shape, int) else tuple(shape)
            shape = shape if len(shape) == 2 else (shape[0], 1)
            self.size, self.shape = int(np.prod(shape)), shape


def entry(name, shape=1, struct=None, expr=None, **kw):
    return Entry(name, shape, struct, expr)


class Struct(object):
    def __init__(self, entries):
        self.entries, self.off = list(entries), {}
        o = 0
        for e in self.entries:
            self.off[e.name] = (o, e)
            o += e.size
        self.size = o

    @property
    def shape(self):
        return (self.size, 1)

    def locate(self, key):
        key = key if isinstance(key, tuple) else (key,)
        o, e = self.off[key[0]]
        if len(key) == 1:
            return o, e.size, e.shape
        o2, n2, sh2 = e.sub.locate(key[1:])
        return o + o2, n2, sh2

    def __call__(self, init=0.):
        return NumStruct(self, init)


class NumStruct(object):
    def __init__(self, layout, init=0.):
        self.layout = layout
        if isinstance(init, NumStruct):
            init = init.cat
        a = np.asarray(init, dtype=float).reshape(-1)
        self.cat = np.full(layout.size, a[0]) if a.size == 1 else a.copy()
        assert self.cat.size == layout.size

    def __getitem__(self, key):
        o, n, sh = self.layout.locate(key)
        return self.cat[o:o + n].reshape(sh, order='F').copy()

    def __setitem__(self, key, value):
        o, n, sh = self.layout.locate(key)
        v = np.asarray(value.a if isinstance(value, mg.MX) else value, dtype=float)
        self.cat[o:o + n] = np.broadcast_to(v.reshape(-1, order='F') if v.size > 1 else v.reshape(-1), (n,))

    def __array__(self, dtype=None, copy=None):
        return self.cat.astype(dtype or float)

    def __deepcopy__(self, memo):
        return NumStruct(self.layout, self.cat.copy())


class SymStruct(object):
    """struct_symMX / struct_MX: only used symbolically by the reference (substitute,
    nlpsol input) -- which the numeric stand-in does not need -- and as a template
    for numeric structs (``constraints(0)``)."""

    def __init__(self, layout):
        self.layout = layout

    def __getitem__(self, key):
        return None

    def __call__(self, init=0.):
        return NumStruct(self.layout, init)


def install_struct_stubs():
    tools = sys.modules['casadi.tools']
    tools.entry = entry
    tools.struct = Struct
    tools.struct_symMX = lambda layout: SymStruct(layout)
    tools.struct_MX = lambda entries: SymStruct(Struct(entries))
    cas = sys.modules['casadi']
    cas.substitute = lambda expr, sym, val: expr


# --------------------------------------------------------------------------
class OracleSolver(object):
    """The solver object behind ``self.problem(x0=, p=, lbg=, ubg=)``: this repository's
    lowered tables + CPU oracle; records what it is given and what it returns."""

    def __init__(self, tables, options=None):
        from oracle import ipm_c
        self.tb, self.ipm_c, self.calls, self.options = tables, ipm_c, [], options

    def __call__(self, x0, p, lbg, ubg, **kw):
        x0, p, lbg, ubg = [np.asarray(v, dtype=float).reshape(-1).copy() for v in (x0, p, lbg, ubg)]
        r = self.ipm_c.solve_batch_full(self.tb, x0[None], p[None], threads=1,
                                        options=self.options, lbg=lbg[None], ubg=ubg[None])
        self.last = r
        self.calls.append((x0, p, lbg, ubg, r['x'][0].copy(), int(r['status'][0])))
        return {'x': r['x'][0], 'lam_g': r['lam_g'][0], 'f': r['f'][0]}

    def stats(self):
        return {'return_status': 'Solve_Succeeded' if self.last['status'][0] == 0
                else 'Restoration_Failed', 'iter_count': int(self.last['iters'][0])}


def run_reference_loop(name, n_steps, update_time, sample_time=0.01, solver_options=None):
    from omg_tools_b200 import scenarios as sc
    tables = getattr(sc, name)(build_solver=False).father.tables
    opt = mg.ref_import('basics.optilayer')
    for cls in list(opt.OptiChild.__subclasses__()) + [opt.OptiChild]:
        if hasattr(cls, '_labels'):
            cls._labels = []
    mg.REG = mg.Registry(seed=3)
    problem = mg.build_reference(name)
    for vehicle in problem.vehicles:            # this framework implements the ideal case
        vehicle.set_options({'ideal_prediction': True, 'ideal_update': True})
        vehicle.problem = problem               # as examples/p2p_3dquadrotor.py:47 does
    solver = OracleSolver(tables, solver_options)
    problem.problem, _ = problem.father.construct_problem(problem.options, problem=solver)
    problem.father.init_transformations(problem.init_primal_transform,
                                        problem.init_dual_transform)
    # Deployer.reset / update (deployer.py:39-79) + Simulator.update (simulator.py:93-99)
    problem.reinitialize()
    t = 0.
    for k in range(n_steps):
        if k == 0:
            problem.initialize(t)
        problem.predict(t, update_time, sample_time, None, None, None, 0, False, False)
        problem.solve(t, update_time)
        problem.store(t, update_time, sample_time)
        problem.simulate(t, update_time, sample_time)
        t = np.round(t + update_time, 6)
    calls = solver.calls
    return {'x0': np.array([c[0] for c in calls]), 'p': np.array([c[1] for c in calls]),
            'lbg': np.array([c[2] for c in calls]), 'ubg': np.array([c[3] for c in calls]),
            'x': np.array([c[4] for c in calls]), 'status': np.array([c[5] for c in calls]),
            'state': np.asarray(problem.vehicles[0].signals['state'], float)[:, -1]}


def main_ext():
    """loop_golden_ext.npz: the same for models added later (default Dubins formulation,
    SimpleQuadrotor3D), through their first knot crossing."""
    mg.install_stubs()
    install_struct_stubs()
    out = {}
    for name, n_steps, dt in (('config_dubins_plain', 6, 0.5), ('config_quadrotor3d_simple', 6, 0.5)):
        res = run_reference_loop(name, n_steps, dt)
        print(name, 'steps', n_steps, 'status', res['status'], 'final state', np.round(res['state'], 4))
        for key, val in res.items():
            out['%s_%s' % (name, key)] = val
        out[name + '_dt'] = dt
    path = OUT.replace('loop_golden.npz', 'loop_golden_ext.npz')
    np.savez_compressed(path, **out)
    print('wrote', path)


def main():
    mg.install_stubs()
    install_struct_stubs()
    out = {}
    for name, n_steps, dt in (('config1', 12, 0.1), ('config5', 12, 0.1), ('config4', 5, 0.4)):
        res = run_reference_loop(name, n_steps, dt)
        print(name, 'steps', n_steps, 'status', res['status'], 'final state', np.round(res['state'], 4))
        for key, val in res.items():
            out['%s_%s' % (name, key)] = val
        out[name + '_dt'] = dt
    # config 4 with the reference's unshifted slacks, the solver's retry option on
    res = run_reference_loop('config4', 13, 0.4, solver_options={'retry_mu': 1e-3})
    print('config4 + retry_mu', 'status', res['status'], 'final state', np.round(res['state'][:3], 4))
    out['config4_retry_status'], out['config4_retry_state'] = res['status'], res['state']
    np.savez_compressed(OUT, **out)
    print('wrote', OUT)


if __name__ == '__main__':
    main_ext() if '--ext' in sys.argv else main()
