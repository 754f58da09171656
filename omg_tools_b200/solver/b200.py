"""ctypes binding of libomgb200.so and the solver object that stands where the
reference holds ``nlpsol('solver','ipopt',...)``.

``B200Solver`` keeps the reference's call contract (problem.py:113-128):

    result = problem(x0=var, p=par, lbg=lb, ubg=ub)   # -> {'x','lam_g','f'}
    problem.stats()['return_status']                  # IPOPT status strings

and adds the batched entry points ``solve_batch`` (host numpy arrays; H2D/D2H
inside the C call) and ``solve_batch_device`` (torch CUDA tensors, zero-copy).
There is no CPU fallback: if the shared library or a CUDA device is missing the
constructor raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'csrc', 'libomgb200.so')

STATUS_STRINGS = {
    0: 'Solve_Succeeded', 1: 'Maximum_Iterations_Exceeded',
    2: 'Restoration_Failed', 3: 'Error_In_Step_Computation',
    4: 'Invalid_Number_Detected', 5: 'Infeasible_Problem_Detected'}

ABI_VERSION = 6
FEAS_STEPS = 30      # option 'feas_steps': LM steps of the feasibility phase (0 = off)

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class _TermList(C.Structure):
    _fields_ = [('n_out', C.c_int32), ('n_terms', C.c_int32), ('width', C.c_int32),
                ('ptr', _i32p), ('coef', _f64p), ('cidx', _i32p), ('xi', _i32p),
                ('lrow', _i32p)]


class _Tables(C.Structure):
    _fields_ = [
        ('abi_version', C.c_int32),
        ('n', C.c_int32), ('m', C.c_int32), ('n_par', C.c_int32),
        ('n_v', C.c_int32), ('degree', C.c_int32),
        ('n_tape', C.c_int32), ('n_tape_terms', C.c_int32), ('n_levels', C.c_int32),
        ('tape_func', _i32p), ('tape_ptr', _i32p), ('tape_coef', _f64p),
        ('tape_fac', _i32p), ('level_ptr', _i32p),
        ('G', _TermList), ('F', _TermList), ('DF', _TermList), ('J', _TermList),
        ('W', _TermList),
        ('nnz_j', C.c_int32), ('jrow', _i32p), ('jcol', _i32p), ('jrow_ptr', _i32p),
        ('n_mid', C.c_int32), ('nnz_jx', C.c_int32), ('n_jp', C.c_int32), ('n_mu', C.c_int32),
        ('jp_ptr', _i32p), ('jp_a', _i32p), ('jp_c', _i32p),
        ('mu_ptr', _i32p), ('mu_row', _i32p), ('mu_slot', _i32p),
        ('nnz_w', C.c_int32), ('wrow', _i32p), ('wcol', _i32p), ('w2h', _i32p),
        ('nnz_h', C.c_int32), ('n_hp', C.c_int32),
        ('hrow', _i32p), ('hcol', _i32p), ('hp_ptr', _i32p),
        ('hp_s1', _i32p), ('hp_s2', _i32p), ('hp_row', _i32p),
        ('lbg', _f64p), ('ubg', _f64p),
        ('kkt_n', C.c_int32), ('kkt_n_eq', C.c_int32), ('env_size', C.c_int32),
        ('n_panel_rows', C.c_int32), ('max_panel_rows', C.c_int32),
        ('kkt_eq_rows', _i32p), ('kkt_pos_var', _i32p), ('kkt_pos_eq', _i32p),
        ('kkt_sign', _i32p), ('env_first', _i32p), ('env_ptr', _i32p),
        ('kkt_hdst', _i32p), ('kkt_jdst', _i32p), ('kkt_diag', _i32p),
        ('kkt_panel_ptr', _i32p), ('kkt_panel_rows', _i32p),
        ('nnz_wx', C.c_int32), ('n_xq', C.c_int32), ('n_xp', C.c_int32),
        ('xq_h', _i32p), ('xq_ptr', _i32p), ('xq_w', _i32p), ('xq_a', _i32p), ('xq_b', _i32p)]


class _Options(C.Structure):
    _fields_ = [('tol', C.c_double), ('constr_viol_tol', C.c_double),
                ('dual_inf_tol', C.c_double), ('compl_inf_tol', C.c_double),
                ('mu_init', C.c_double), ('bound_push', C.c_double),
                ('bound_frac', C.c_double), ('mult_bound_push', C.c_double),
                ('bound_relax_factor', C.c_double),
                ('scaling_max_gradient', C.c_double),
                ('max_iter', C.c_int32), ('trace', C.c_int32),
                ('max_restarts', C.c_int32), ('soft_resto', C.c_int32),
                ('restart_mu', C.c_double), ('restart_push', C.c_double),
                ('inertia_mode', C.c_int32), ('reserved', C.c_int32)]


EXPORTS = ['omg_abi_version', 'omg_last_error', 'omg_default_options',
           'omg_problem_create', 'omg_problem_destroy', 'omg_set_options',
           'omg_solve_batch', 'omg_solve_batch_host', 'omg_shift_batch',
           'omg_get_trace', 'omg_get_info', 'omg_structure_info', 'omg_last_timing',
           'omg_admm_zl_update', 'omg_sample_batch', 'omg_tables_read',
           'omg_tables_free', 'omg_integrate_rk4', 'omg_feas_batch', 'omg_feas_batch_host',
           'omg_comm_unique_id', 'omg_comm_create', 'omg_comm_destroy', 'omg_admm_exchange_x',
           'omg_admm_zl_update_dist']

_lib = None


def load_library():
    """Load libomgb200.so (built by __graft_entry__.build() / csrc/Makefile).  There is exactly
    one library and no fallback: a missing build raises."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libomgb200.so not found at %s: build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` '
            '(there is no CPU fallback)' % LIB_PATH)
    _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def bind(lib):
    """Declare the C signatures of include/omg_b200.h on a loaded library object."""
    lib.omg_abi_version.restype = C.c_int
    lib.omg_last_error.restype = C.c_char_p
    lib.omg_default_options.argtypes = [C.POINTER(_Options)]
    lib.omg_default_options.restype = None
    lib.omg_problem_create.argtypes = [C.POINTER(_Tables), C.POINTER(_Options), C.c_int]
    lib.omg_problem_create.restype = C.c_void_p
    lib.omg_problem_destroy.argtypes = [C.c_void_p]
    lib.omg_problem_destroy.restype = None
    lib.omg_set_options.argtypes = [C.c_void_p, C.POINTER(_Options)]
    vp = C.c_void_p
    lib.omg_solve_batch.argtypes = [vp, C.c_int32, vp, vp, vp, vp, C.c_int32, vp,
                                    vp, vp, vp, vp, vp, vp]
    lib.omg_solve_batch_host.argtypes = [vp, C.c_int32, vp, vp, vp, vp, C.c_int32,
                                         vp, vp, vp, vp, vp, vp]
    lib.omg_shift_batch.argtypes = [vp, C.c_int32, vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.omg_feas_batch.argtypes = [vp, C.c_int32, vp, vp, vp, vp, C.c_int32, C.c_int32, vp, vp, vp, vp]
    lib.omg_feas_batch_host.argtypes = [vp, C.c_int32, vp, vp, vp, vp, C.c_int32, C.c_int32, vp, vp, vp]
    lib.omg_get_trace.argtypes = [vp, vp, C.c_int32]
    lib.omg_get_info.argtypes = [vp] + [_i32p] * 6
    lib.omg_structure_info.argtypes = [vp]
    lib.omg_structure_info.restype = C.c_char_p
    lib.omg_comm_unique_id.argtypes = [vp]
    lib.omg_comm_create.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    lib.omg_comm_create.restype = C.c_void_p
    lib.omg_comm_destroy.argtypes = [vp]
    lib.omg_comm_destroy.restype = None
    lib.omg_admm_exchange_x.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp]
    lib.omg_admm_zl_update_dist.argtypes = [vp] + [C.c_int32] * 4 + [vp] * 4 + [C.c_double] + [vp] * 13
    lib.omg_last_timing.argtypes = [vp, C.POINTER(C.c_float), _i32p]
    lib.omg_admm_zl_update.argtypes = [C.c_int32] * 4 + [vp] * 4 + [C.c_double] + [vp] * 8
    lib.omg_sample_batch.argtypes = [C.c_int32, C.c_int32, vp, C.c_int32] + [vp] * 7
    lib.omg_integrate_rk4.argtypes = [C.c_int32] * 4 + [vp, vp, C.c_double, C.c_int32, vp, vp]
    lib.omg_tables_read.argtypes = [C.c_char_p]
    lib.omg_tables_read.restype = C.POINTER(_Tables)
    lib.omg_tables_free.argtypes = [C.POINTER(_Tables)]
    lib.omg_tables_free.restype = None
    return lib


def _ptr(arr, typ):
    return arr.ctypes.data_as(typ)


def _check_device_tensors(tensors, lib=None):
    """Device-pointer API: contiguous float64 CUDA tensors, nothing else."""
    import torch
    for t in tensors:
        if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
            raise ValueError('expected contiguous float64 CUDA tensors')
    return True


def _check_int_tensors(tensors):
    import torch
    for t in tensors:
        if t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous():
            raise ValueError('expected contiguous int32 CUDA tensors')


def _stream_handle(on_gpu, device, stream):
    import torch
    if not on_gpu:
        return None
    if stream is None:
        stream = torch.cuda.current_stream(device)
    return C.c_void_p(stream.cuda_stream)


class _Keep(object):
    """Owns contiguous numpy copies referenced by a ctypes struct."""

    def __init__(self):
        self.arrays = []

    def i32(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        self.arrays.append(a)
        return _ptr(a, _i32p)

    def f64(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self.arrays.append(a)
        return _ptr(a, _f64p)


def pack_tables(tb):
    """NLPTables -> (ctypes omg_tables, keep-alive object)."""
    keep = _Keep()

    def tl(t, with_lrow=False):
        s = _TermList()
        s.n_out, s.n_terms, s.width = t.n_out, t.n_terms, t.width
        s.ptr, s.coef = keep.i32(t.ptr), keep.f64(t.coef)
        s.cidx, s.xi = keep.i32(t.cidx), keep.i32(t.xi.reshape(-1))
        s.lrow = keep.i32(t.lrow) if with_lrow else None
        return s

    T = _Tables()
    T.abi_version = ABI_VERSION
    T.n, T.m, T.n_par, T.n_v, T.degree = tb.n, tb.m, tb.n_par, tb.n_v, tb.degree
    T.n_tape, T.n_tape_terms = len(tb.tape_func), len(tb.tape_coef)
    T.n_levels = len(tb.level_ptr) - 1
    T.tape_func, T.tape_ptr = keep.i32(tb.tape_func), keep.i32(tb.tape_ptr)
    T.tape_coef, T.tape_fac = keep.f64(tb.tape_coef), keep.i32(tb.tape_fac.reshape(-1))
    T.level_ptr = keep.i32(tb.level_ptr)
    T.G, T.F, T.DF, T.J = tl(tb.G), tl(tb.F), tl(tb.DF), tl(tb.J)
    T.W = tl(tb.W, True)
    T.nnz_j = tb.nnz_j
    T.jrow, T.jcol, T.jrow_ptr = keep.i32(tb.jrow), keep.i32(tb.jcol), keep.i32(tb.jrow_ptr)
    T.n_mid = getattr(tb, 'n_mid', 0)
    T.nnz_jx = getattr(tb, 'nnz_jx', tb.nnz_j)
    if T.n_mid:
        T.n_jp, T.n_mu = len(tb.jp_a), len(tb.mu_row)
        T.jp_ptr, T.jp_a, T.jp_c = keep.i32(tb.jp_ptr), keep.i32(tb.jp_a), keep.i32(tb.jp_c)
        T.mu_ptr, T.mu_row, T.mu_slot = (keep.i32(tb.mu_ptr), keep.i32(tb.mu_row),
                                         keep.i32(tb.mu_slot))
    T.nnz_w = tb.nnz_w
    T.wrow, T.wcol, T.w2h = keep.i32(tb.wrow), keep.i32(tb.wcol), keep.i32(tb.w2h)
    T.nnz_h, T.n_hp = tb.nnz_h, len(tb.hp_s1)
    T.hrow, T.hcol, T.hp_ptr = keep.i32(tb.hrow), keep.i32(tb.hcol), keep.i32(tb.hp_ptr)
    T.hp_s1, T.hp_s2, T.hp_row = keep.i32(tb.hp_s1), keep.i32(tb.hp_s2), keep.i32(tb.hp_row)
    T.lbg, T.ubg = keep.f64(tb.lbg), keep.f64(tb.ubg)
    T.kkt_n, T.kkt_n_eq, T.env_size = tb.kkt_n, tb.kkt_n_eq, tb.env_size
    T.n_panel_rows, T.max_panel_rows = len(tb.kkt_panel_rows), tb.kkt_max_panel_rows
    T.kkt_eq_rows, T.kkt_pos_var = keep.i32(tb.kkt_eq_rows), keep.i32(tb.kkt_pos_var)
    T.kkt_pos_eq, T.kkt_sign = keep.i32(tb.kkt_pos_eq), keep.i32(tb.kkt_sign)
    T.env_first, T.env_ptr = keep.i32(tb.env_first), keep.i32(tb.env_ptr)
    T.kkt_hdst, T.kkt_jdst = keep.i32(tb.kkt_hdst), keep.i32(tb.kkt_jdst)
    T.kkt_diag = keep.i32(tb.kkt_diag)
    T.kkt_panel_ptr = keep.i32(tb.kkt_panel_ptr)
    T.kkt_panel_rows = keep.i32(tb.kkt_panel_rows)
    T.nnz_wx = getattr(tb, 'nnz_wx', 0)
    if T.nnz_wx:
        T.n_xq, T.n_xp = tb.n_xq, len(tb.xq_w)
        T.xq_h, T.xq_ptr = keep.i32(tb.xq_h), keep.i32(tb.xq_ptr)
        T.xq_w, T.xq_a, T.xq_b = keep.i32(tb.xq_w), keep.i32(tb.xq_a), keep.i32(tb.xq_b)
    return T, keep


_NO_EFFECT_IPOPT_OPTIONS = frozenset([
    'print_level', 'print_time', 'sb', 'file_print_level', 'output_file',
    'print_timing_statistics', 'print_user_options', 'linear_solver',
    'warm_start_init_point', 'fixed_variable_treatment', 'ma57_automatic_scaling',
    'hessian_approximation', 'verbose'])

_TERMLIST_FIELDS = ('G', 'F', 'DF', 'J', 'W')


def _table_records(tb):
    """[(name, dtype, array)] in the order of include/omg_b200.h; dtype 0 =
    int32, 1 = float64; scalars are int32 arrays of length 1."""
    T, keep = pack_tables(tb)
    rec = []

    def scalar(name, v):
        rec.append((name, 0, np.array([v], dtype=np.int32)))

    def arr(name, a, dtype):
        a = np.ascontiguousarray(a, dtype=np.float64 if dtype else np.int32).reshape(-1)
        rec.append((name, dtype, a))

    for f in ('n', 'm', 'n_par', 'n_v', 'degree', 'n_tape', 'n_tape_terms', 'n_levels'):
        scalar(f, getattr(T, f))
    arr('tape_func', tb.tape_func, 0), arr('tape_ptr', tb.tape_ptr, 0)
    arr('tape_coef', tb.tape_coef, 1), arr('tape_fac', tb.tape_fac, 0)
    arr('level_ptr', tb.level_ptr, 0)
    for l in _TERMLIST_FIELDS:
        t = getattr(tb, l)
        scalar(l + '.n_out', t.n_out), scalar(l + '.n_terms', t.n_terms)
        scalar(l + '.width', t.width)
        arr(l + '.ptr', t.ptr, 0), arr(l + '.coef', t.coef, 1), arr(l + '.cidx', t.cidx, 0)
        arr(l + '.xi', t.xi, 0)
        if l == 'W':
            arr(l + '.lrow', t.lrow, 0)
    scalar('nnz_j', tb.nnz_j)
    arr('jrow', tb.jrow, 0), arr('jcol', tb.jcol, 0), arr('jrow_ptr', tb.jrow_ptr, 0)
    n_mid = getattr(tb, 'n_mid', 0)
    scalar('n_mid', n_mid), scalar('nnz_jx', getattr(tb, 'nnz_jx', tb.nnz_j))
    scalar('n_jp', len(tb.jp_a) if n_mid else 0), scalar('n_mu', len(tb.mu_row) if n_mid else 0)
    empty = np.zeros(0, dtype=np.int32)
    for f in ('jp_ptr', 'jp_a', 'jp_c', 'mu_ptr', 'mu_row', 'mu_slot'):
        arr(f, getattr(tb, f) if n_mid else empty, 0)
    scalar('nnz_w', tb.nnz_w)
    arr('wrow', tb.wrow, 0), arr('wcol', tb.wcol, 0), arr('w2h', tb.w2h, 0)
    scalar('nnz_h', tb.nnz_h), scalar('n_hp', len(tb.hp_s1))
    for f in ('hrow', 'hcol', 'hp_ptr', 'hp_s1', 'hp_s2', 'hp_row'):
        arr(f, getattr(tb, f), 0)
    arr('lbg', tb.lbg, 1), arr('ubg', tb.ubg, 1)
    scalar('kkt_n', tb.kkt_n), scalar('kkt_n_eq', tb.kkt_n_eq), scalar('env_size', tb.env_size)
    scalar('n_panel_rows', len(tb.kkt_panel_rows)), scalar('max_panel_rows', tb.kkt_max_panel_rows)
    for f in ('kkt_eq_rows', 'kkt_pos_var', 'kkt_pos_eq', 'kkt_sign', 'env_first', 'env_ptr',
              'kkt_hdst', 'kkt_jdst', 'kkt_diag', 'kkt_panel_ptr', 'kkt_panel_rows'):
        arr(f, getattr(tb, f), 0)
    nnz_wx = getattr(tb, 'nnz_wx', 0)
    scalar('nnz_wx', nnz_wx), scalar('n_xq', tb.n_xq if nnz_wx else 0)
    scalar('n_xp', len(tb.xq_w) if nnz_wx else 0)
    for f in ('xq_h', 'xq_ptr', 'xq_w', 'xq_a', 'xq_b'):
        arr(f, getattr(tb, f) if nnz_wx else empty, 0)
    del keep
    return rec


def save_tables(tb, path):
    """Write the lowered NLP to a table file (include/omg_b200.h:
    omg_tables_read) -- the deployable artefact for native callers, in place
    of the nlp.c/.so bundle of the reference's exporter."""
    import struct
    rec = _table_records(tb)
    with open(path, 'wb') as fp:
        fp.write(b'OMGTBL\0\0')
        fp.write(struct.pack('<ii', ABI_VERSION, len(rec)))
        for name, dtype, a in rec:
            nb = name.encode()
            if len(nb) > 23:
                raise ValueError('record name too long: %s' % name)
            fp.write(nb.ljust(24, b'\0'))
            fp.write(struct.pack('<iiq', dtype, 0, a.size))
            fp.write(a.tobytes())


class B200Solver(object):
    """Batched interior-point solver on one B200 for one NLP structure."""

    def __init__(self, tables, options=None, device=None):
        self.lib = load_library()
        if self.lib.omg_abi_version() != ABI_VERSION:
            raise RuntimeError('libomgb200.so ABI version mismatch')
        self.tables = tables
        self.n, self.m, self.n_par = tables.n, tables.m, tables.n_par
        self._opt = _Options()
        self.lib.omg_default_options(C.byref(self._opt))
        self._apply_options(options or {})
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', '0')) \
                if 'OMG_B200_DEVICE' not in os.environ \
                else int(os.environ['OMG_B200_DEVICE'])
        self.device = device
        T, keep = pack_tables(tables)
        self._handle = self.lib.omg_problem_create(C.byref(T), C.byref(self._opt),
                                                   int(device))
        del keep
        if not self._handle:
            raise RuntimeError('omg_problem_create failed: %s' %
                               self.lib.omg_last_error().decode())
        self._stats = {'return_status': None, 'iter_count': 0}

    def _apply_options(self, options):
        import warnings
        fields = dict((f[0], f[1]) for f in _Options._fields_)
        for key, value in options.items():
            if key in fields and key != 'reserved':
                setattr(self._opt, key, value)
            elif key == 'retry_mu':
                self._retry_mu = float(value)
            elif key == 'feas_steps':
                self._feas_steps = int(value)
            elif key in _NO_EFFECT_IPOPT_OPTIONS:
                # printing / linear-solver selection, and warm_start_init_point: the
                # warm-start pushes are always the 'yes' variants the reference sets
                pass
            else:
                # an IPOPT option that changes the algorithm would silently be lost
                warnings.warn('solver option %r is not supported by the b200 solver and '
                              'is ignored' % key)

    def set_options(self, options):
        self._apply_options(options)
        self._check(self.lib.omg_set_options(self._handle, C.byref(self._opt)))

    def __del__(self):
        try:
            if getattr(self, '_handle', None):
                self.lib.omg_problem_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError('libomgb200: %s' % self.lib.omg_last_error().decode())

    # ------------------------------------------------------------------
    def info(self):
        vals = [C.c_int32() for _ in range(6)]
        self._check(self.lib.omg_get_info(self._handle, *[C.byref(v) for v in vals]))
        keys = ('n', 'm', 'n_par', 'smem_bytes', 'ctas_per_sm', 'n_sm')
        return dict(zip(keys, [v.value for v in vals]))

    @property
    def structure(self):
        """One-line report of the kernel family / factor structure chosen for this problem."""
        return self.lib.omg_structure_info(self._handle).decode()

    def last_timing(self):
        ms, nl = C.c_float(), C.c_int32()
        self._check(self.lib.omg_last_timing(self._handle, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def trace(self, max_rows=512):
        out = np.zeros((max_rows, 8))
        rows = self.lib.omg_get_trace(self._handle, out.ctypes.data, max_rows)
        if rows < 0:
            self._check(rows)
        return out[:rows]

    # ------------------------------------------------------------------
    def _bounds(self, lbg, ubg, B):
        lbg = self.tables.lbg if lbg is None else np.asarray(lbg, dtype=np.float64)
        ubg = self.tables.ubg if ubg is None else np.asarray(ubg, dtype=np.float64)
        lbg = np.ascontiguousarray(np.asarray(lbg, dtype=np.float64))
        ubg = np.ascontiguousarray(np.asarray(ubg, dtype=np.float64))
        shared = 1 if lbg.ndim == 1 else 0
        if (shared and lbg.size != self.m) or (not shared and lbg.shape != (B, self.m)):
            raise ValueError('lbg/ubg must have shape (m,) or (B, m)')
        if ubg.shape != lbg.shape:
            raise ValueError('lbg and ubg shapes differ')
        return lbg, ubg, shared

    def solve_batch(self, X0, P, lbg=None, ubg=None, lam_g0=None, _retry=True):
        """Host arrays in, host arrays out (H2D + solve + D2H in one C call).

        Instances that end in Restoration_Failed go through the feasibility phase
        (option ``feas_steps``, default 30 Levenberg-Marquardt steps, 0 = off) and are
        solved once more from the point it returns; their iteration counts add up.

        Option ``retry_mu`` > 0 (default 0 = off): instances that did not succeed are
        solved once more from the same start with ``mu_init = retry_mu`` (e.g. 1e-3).
        A small initial barrier parameter keeps the iterates near an infeasible warm
        start instead of pushing every slack to the centre first; it rescues the warm
        starts that IPOPT leaves to its restoration phase."""
        X0 = np.ascontiguousarray(X0, dtype=np.float64).reshape(-1, self.n)
        B = X0.shape[0]
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(B, self.n_par)
        lbg, ubg, shared = self._bounds(lbg, ubg, B)
        lam0 = None
        if lam_g0 is not None:
            lam0 = np.ascontiguousarray(lam_g0, dtype=np.float64).reshape(B, self.m)
        X = np.empty((B, self.n))
        LAM = np.empty((B, self.m))
        F = np.empty(B)
        status = np.empty(B, dtype=np.int32)
        iters = np.empty(B, dtype=np.int32)
        self._check(self.lib.omg_solve_batch_host(
            self._handle, B, X0.ctypes.data, P.ctypes.data, lbg.ctypes.data,
            ubg.ctypes.data, shared, lam0.ctypes.data if lam0 is not None else None,
            X.ctypes.data, LAM.ctypes.data, F.ctypes.data, status.ctypes.data,
            iters.ctypes.data))
        res = {'x': X, 'lam_g': LAM, 'f': F, 'status': status, 'iters': iters}
        n_feas = getattr(self, '_feas_steps', FEAS_STEPS)
        if _retry and n_feas > 0 and (status == 2).any():
            # Restoration_Failed: feasibility phase from the point where the line search
            # gave up, then one more solve from there
            idx = np.nonzero(status == 2)[0]
            lb_i, ub_i = (lbg, ubg) if shared else (lbg[idx], ubg[idx])
            x1, _, _ = self.feasibility_batch(X[idx], P[idx], lb_i, ub_i, n_feas)
            r2 = self.solve_batch(x1, P[idx], lb_i, ub_i, None, _retry=False)
            ok = r2['status'] == 0          # the first result stands unless the re-solve succeeds
            for key in ('x', 'lam_g', 'f', 'status'):
                res[key][idx[ok]] = r2[key][ok]
            res['iters'][idx] += r2['iters']
        mu_r = getattr(self, '_retry_mu', 0.)
        if _retry and mu_r > 0. and (status != 0).any():
            idx = np.nonzero(status != 0)[0]
            mu_old = self._opt.mu_init
            self.set_options({'mu_init': mu_r})
            try:
                r2 = self.solve_batch(X0[idx], P[idx], lbg if shared else lbg[idx],
                                      ubg if shared else ubg[idx],
                                      None if lam0 is None else lam0[idx], _retry=False)
            finally:
                self.set_options({'mu_init': mu_old})
            ok = r2['status'] == 0
            for key in ('x', 'lam_g', 'f', 'status'):
                res[key][idx[ok]] = r2[key][ok]
            res['iters'][idx] += r2['iters']
        return res

    def feasibility_batch(self, X0, P, lbg=None, ubg=None, max_steps=None):
        """Host arrays in / out: up to ``max_steps`` Levenberg-Marquardt steps on the
        constraint violation from X0 (omg_feas_batch_host).  Returns (X, max |violation|
        per instance, steps taken per instance)."""
        X0 = np.ascontiguousarray(X0, dtype=np.float64).reshape(-1, self.n)
        B = X0.shape[0]
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(B, self.n_par)
        lbg, ubg, shared = self._bounds(lbg, ubg, B)
        if max_steps is None:
            max_steps = getattr(self, '_feas_steps', FEAS_STEPS)
        X = np.empty((B, self.n))
        viol = np.empty(B)
        steps = np.empty(B, dtype=np.int32)
        self._check(self.lib.omg_feas_batch_host(
            self._handle, B, X0.ctypes.data, P.ctypes.data, lbg.ctypes.data, ubg.ctypes.data,
            shared, int(max_steps), X.ctypes.data, viol.ctypes.data, steps.ctypes.data))
        return X, viol, steps

    def solve_batch_device(self, X0, P, LBG, UBG, X, LAM, F, STATUS, ITERS,
                           lam_g0=None, stream=None):
        """torch CUDA tensors (float64 / int32, contiguous); asynchronous on
        ``stream`` (torch.cuda.Stream or None = current)."""
        B = X0.shape[0]
        on_gpu = _check_device_tensors((X0, P, LBG, UBG, X, LAM, F) +
                                       ((lam_g0,) if lam_g0 is not None else ()), self.lib)
        _check_int_tensors((STATUS, ITERS))
        if (X0.shape != (B, self.n) or P.shape != (B, self.n_par) or X.shape != (B, self.n) or
                LAM.shape != (B, self.m) or F.numel() != B or STATUS.numel() != B or ITERS.numel() != B or
                (lam_g0 is not None and lam_g0.shape != (B, self.m))):
            raise ValueError('tensor shapes do not match the batch / problem sizes')
        shared = 1 if LBG.dim() == 1 else 0
        self._check(self.lib.omg_solve_batch(
            self._handle, B, X0.data_ptr(), P.data_ptr(), LBG.data_ptr(),
            UBG.data_ptr(), shared, lam_g0.data_ptr() if lam_g0 is not None else None,
            X.data_ptr(), LAM.data_ptr(), F.data_ptr(), STATUS.data_ptr(),
            ITERS.data_ptr(), _stream_handle(on_gpu, X0.device, stream)))

    def shift_batch_device(self, X, blocks, stream=None):
        """In-place warm-start shift of spline variables (torch CUDA tensor X
        [B, n]); blocks = [(offset, len_basis, n_columns, T)]."""
        import torch
        offs = np.array([b[0] for b in blocks], dtype=np.int32)
        lens = np.array([b[1] for b in blocks], dtype=np.int32)
        ncols = np.array([b[2] for b in blocks], dtype=np.int32)
        Tm = np.concatenate([np.asarray(b[3], dtype=np.float64).reshape(-1)
                             for b in blocks])
        on_gpu = _check_device_tensors((X,), self.lib)
        self._check(self.lib.omg_shift_batch(
            self._handle, X.shape[0], X.data_ptr(), len(blocks), offs.ctypes.data,
            lens.ctypes.data, ncols.ctypes.data, Tm.ctypes.data,
            _stream_handle(on_gpu, X.device, stream)))

    # ------------------------------------------------------------------
    # the reference's single-instance call contract
    # ------------------------------------------------------------------
    def __call__(self, x0=None, p=None, lbg=None, ubg=None, lam_g0=None, **kwargs):
        x0 = np.asarray(x0, dtype=np.float64).reshape(1, self.n)
        p = np.asarray(p, dtype=np.float64).reshape(1, self.n_par)
        lb = None if lbg is None else np.asarray(lbg, dtype=np.float64).reshape(-1)
        ub = None if ubg is None else np.asarray(ubg, dtype=np.float64).reshape(-1)
        lam0 = None if lam_g0 is None else \
            np.asarray(lam_g0, dtype=np.float64).reshape(1, self.m)
        if lb is not None and ub is not None:
            # the equality rows are part of the factorised structure: bounds that turn an equality
            # into an inequality (or the reverse) need a re-lowered problem, not another call
            eq_now, eq_built = (lb == ub), (self.tables.lbg == self.tables.ubg)
            if not np.array_equal(eq_now, eq_built):
                rows = np.nonzero(eq_now != eq_built)[0][:5].tolist()
                raise ValueError('lbg/ubg change which rows are equalities (rows %s ...): the '
                                 'equality pattern is fixed when the problem is lowered' % rows)
        res = self.solve_batch(x0, p, lb, ub, lam0)
        self._stats = {'return_status': STATUS_STRINGS[int(res['status'][0])],
                       'iter_count': int(res['iters'][0]),
                       'success': int(res['status'][0]) == 0}
        return {'x': res['x'][0], 'lam_g': res['lam_g'][0], 'f': float(res['f'][0])}

    def stats(self):
        return dict(self._stats)


def admm_zl_update(PzT, c, Tf, Tb, rho, x_i, x_j, z_i, z_ij, l_i, l_ij, res, L, stream=None):
    """Consensus step (z, lambda, residuals) for the agents held in the given
    torch CUDA tensors; z_*, l_* are updated in place (omg_admm_zl_update)."""
    import torch
    lib = load_library()
    n_agents, nsh = x_i.shape
    nn = x_j.shape[1]
    on_gpu = _check_device_tensors((PzT, c, Tf, Tb, x_i, x_j, z_i, z_ij, l_i, l_ij, res), lib)
    rc = lib.omg_admm_zl_update(n_agents, nsh, nn, L, PzT.data_ptr(), c.data_ptr(),
                                Tf.data_ptr(), Tb.data_ptr(), float(rho), x_i.data_ptr(),
                                x_j.data_ptr(), z_i.data_ptr(), z_ij.data_ptr(),
                                l_i.data_ptr(), l_ij.data_ptr(), res.data_ptr(),
                                _stream_handle(on_gpu, x_i.device, stream))
    if rc != 0:
        raise RuntimeError('libomgb200: %s' % lib.omg_last_error().decode())


class AdmmComm(object):
    """The library's NCCL communicator for the formation ADMM exchange (include/omg_b200.h:
    omg_comm_*).  The 128-byte unique id is created on rank 0 and handed to the other ranks by
    the caller's bootstrap -- here a torch.distributed broadcast; a C++ caller uses whatever it
    has (MPI, a socket, a file)."""

    def __init__(self, rank=0, world=1, device=0, group=None):
        import torch
        self.lib = load_library()
        idbuf = (C.c_char * 128)()
        if world > 1:
            import torch.distributed as dist
            t = torch.zeros(128, dtype=torch.uint8, device=torch.device('cuda', device))
            if rank == 0:
                if self.lib.omg_comm_unique_id(idbuf) != 0:
                    raise RuntimeError('libomgb200: %s' % self.lib.omg_last_error().decode())
                t.copy_(torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8))
            dist.broadcast(t, src=0, group=group)
            idbuf.raw = bytes(t.cpu().numpy().tobytes())
        self.handle = self.lib.omg_comm_create(idbuf, world, rank, device)
        if not self.handle:
            raise RuntimeError('libomgb200: %s' % self.lib.omg_last_error().decode())
        self.rank, self.world, self.device = rank, world, device

    def exchange_x(self, nghb, x_i, x_j, stream=None):
        n_local, nsh = x_i.shape
        rc = self.lib.omg_admm_exchange_x(self.handle, n_local, nsh, x_j.shape[1], nghb.data_ptr(),
                                          x_i.data_ptr(), x_j.data_ptr(), _stream_handle(True, x_i.device, stream))
        if rc != 0:
            raise RuntimeError('libomgb200: %s' % self.lib.omg_last_error().decode())

    def zl_update(self, PzT, c, Tf, Tb, rho, x_i, x_j, z_i, z_ij, l_i, l_ij, res, L, nghb, back,
                  z_ji, l_ji, res_total, stream=None):
        """Consensus kernel + residual all-reduce + second exchange, one stream-ordered call."""
        n_local, nsh = x_i.shape
        rc = self.lib.omg_admm_zl_update_dist(
            self.handle, n_local, nsh, x_j.shape[1], L, PzT.data_ptr(), c.data_ptr(), Tf.data_ptr(),
            Tb.data_ptr(), float(rho), x_i.data_ptr(), x_j.data_ptr(), z_i.data_ptr(), z_ij.data_ptr(),
            l_i.data_ptr(), l_ij.data_ptr(), res.data_ptr(), nghb.data_ptr(), back.data_ptr(),
            z_ji.data_ptr(), l_ji.data_ptr(), res_total.data_ptr(), _stream_handle(True, x_i.device, stream))
        if rc != 0:
            raise RuntimeError('libomgb200: %s' % self.lib.omg_last_error().decode())

    def __del__(self):
        try:
            if self.handle:
                self.lib.omg_comm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def sample_batch(X, blocks, stream=None):
    """Apply sampling matrices to spline variables of a batch on the device.
    X: torch CUDA tensor [B, n]; blocks = [(offset, len_basis, n_columns, S[nsamp, len])].
    Returns a CUDA tensor [B, sum(nsamp * n_columns)] laid out block / column / sample."""
    import torch
    lib = load_library()
    offs = np.array([b[0] for b in blocks], dtype=np.int32)
    lens = np.array([b[1] for b in blocks], dtype=np.int32)
    ncols = np.array([b[2] for b in blocks], dtype=np.int32)
    nsamp = np.array([np.asarray(b[3]).shape[0] for b in blocks], dtype=np.int32)
    Sm = np.concatenate([np.ascontiguousarray(b[3], dtype=np.float64).reshape(-1) for b in blocks])
    out = torch.empty((X.shape[0], int((nsamp * ncols).sum())), dtype=torch.float64, device=X.device)
    on_gpu = _check_device_tensors((X,), lib)
    rc = lib.omg_sample_batch(X.shape[0], X.shape[1], X.data_ptr(), len(blocks), offs.ctypes.data,
                              lens.ctypes.data, ncols.ctypes.data, nsamp.ctypes.data,
                              Sm.ctypes.data, out.data_ptr(), _stream_handle(on_gpu, X.device, stream))
    if rc != 0:
        raise RuntimeError('libomgb200: %s' % lib.omg_last_error().decode())
    return out


ODE_MODELS = {'Holonomic': 0, 'Holonomic1D': 0, 'Holonomic3D': 0, 'Quadrotor3D': 1, 'Quadrotor': 2}


def integrate_rk4(model, state0, inputs, sample_time, stream=None):
    """Non-ideal prediction on the device (omg_integrate_rk4): state0 [B, n_state] and the
    planned inputs [B, steps+1, n_input] are torch CUDA float64 tensors; returns the states
    after steps*sample_time.  ``model`` is a vehicle class name or a model id."""
    import torch
    lib = load_library()
    mid = ODE_MODELS[model] if isinstance(model, str) else int(model)
    on_gpu = _check_device_tensors((state0, inputs), lib)
    B, ns = state0.shape
    steps, ni = inputs.shape[1] - 1, inputs.shape[2]
    out = torch.empty_like(state0)
    rc = lib.omg_integrate_rk4(mid, B, ns, ni, state0.data_ptr(), inputs.data_ptr(),
                               float(sample_time), steps, out.data_ptr(),
                               _stream_handle(on_gpu, state0.device, stream))
    if rc != 0:
        raise RuntimeError('libomgb200: %s' % lib.omg_last_error().decode())
    return out
