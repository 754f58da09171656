"""CasADi expression graph -> omg_tables: the binding that lets the REFERENCE's own model feed
the B200 solver.

The reference builds its NLP as CasADi ``MX`` expressions and hands ``{x, p, f, g}`` to
``nlpsol`` in ``create_nlp`` (omgtools/basics/optilayer.py:49-60; the symbols come from
``construct_variables / construct_parameters / construct_constraints / construct_objective``,
optilayer.py:225-279).  ``lower_casadi(var, par, obj, con, lbg, ubg)`` takes exactly those four
expressions and returns the lowered tables ``B200Solver`` consumes, so that the one branch a
maintainer adds to ``create_nlp`` is

    if options['solver'] == 'b200':
        from omg_tools_b200.basics.lower_casadi import lower_casadi
        from omg_tools_b200.solver.b200 import B200Solver
        tables = lower_casadi(var, par, obj, con, lbg, ubg)
        return B200Solver(tables, options['solver_options'].get('b200', {})), buildtime

How: ``Function('nlp', [x, p], [f, g]).expand()`` turns the MX graph into CasADi's scalar (SX)
virtual machine, whose instruction list is public API (``n_instructions / instruction_id /
instruction_input / instruction_output / instruction_constant``; CasADi's example
"accessing_sx_algorithm").  The list is interpreted once with THIS framework's symbolic scalar
(basics/poly.py) in the work vector: additions and products of polynomials in x with
parameter-only coefficients, and the non-polynomial operations -- division, comparisons (the
Cox-de Boor indicators of ``evalspline``), sin / cos / sqrt -- applied to parameter-only values
become atoms of the parameter tape.  The result is one ``Poly`` per constraint row and one for
the objective: the input of ``lowering.lower``, as if the model had been written with this
framework's own modelling classes.

CasADi is not installed in the authoring image.  ``lower_sx_function`` therefore takes the
instruction-level interface as a duck-typed object plus the table of operation codes, and the
test suite drives it with a recording of the graph the REFERENCE's modelling code builds
(tests/golden/make_casadi_graph_golden.py, tests/test_lower_casadi.py).
"""
import numpy as np

from . import poly
from .lowering import lower
from .poly import Poly

# the operation names of casadi (casadi.OP_*) this interpreter understands
_BINARY = {
    'OP_ADD': lambda a, b: a + b,
    'OP_SUB': lambda a, b: a - b,
    'OP_MUL': lambda a, b: a * b,
    'OP_DIV': lambda a, b: _div(a, b),
    'OP_LT': lambda a, b: _cmp(a, b, strict=True),      # a < b
    'OP_LE': lambda a, b: _cmp(a, b, strict=False),     # a <= b
    'OP_AND': lambda a, b: a * b,                       # product of 0/1 indicators
}
_UNARY = {
    'OP_ASSIGN': lambda a: a,
    'OP_NEG': lambda a: -a,
    'OP_SQ': lambda a: a * a,
    'OP_TWICE': lambda a: 2.0 * a,
    'OP_INV': lambda a: _div(1.0, a),
    'OP_SIN': lambda a: poly.sin(a),
    'OP_COS': lambda a: poly.cos(a),
    'OP_SQRT': lambda a: poly.sqrt(a),
}


def _as_poly(a):
    return a if isinstance(a, Poly) else Poly({(): float(a)} if float(a) != 0.0 else {})


def _div(a, b):
    if not isinstance(b, Poly):
        return a * (1.0 / float(b))
    return _as_poly(a) / b              # parameter-only denominator -> 'inv' atom (poly.py)


def _cmp(a, b, strict):
    a, b = _as_poly(a), _as_poly(b)
    return (a < b) if strict else (a <= b)   # 0/1 indicator atoms of parameter-only arguments


def op_table(module):
    """{code: name} for the operation codes of a casadi-like module (casadi itself, or the
    stand-in of the tests)."""
    names = list(_BINARY) + list(_UNARY) + ['OP_CONST', 'OP_INPUT', 'OP_OUTPUT', 'OP_CONSTPOW', 'OP_POW',
                                           'OP_IF_ELSE_ZERO', 'OP_NOT', 'OP_FABS']
    return {getattr(module, nm): nm for nm in names if hasattr(module, nm)}


def lower_sx_function(f, lbg, ubg, ops, order_hint=None, names=None):
    """Interpret the instruction list of an SX function (inputs: x, p; outputs: f, g) with
    polynomials and lower the rows.  ``f`` needs: n_instructions(), instruction_id(k),
    instruction_input(k), instruction_output(k), instruction_constant(k), sz_w(), nnz_in(i),
    nnz_out(i).  ``ops``: {operation code: 'OP_*' name} (op_table)."""
    n, n_par, m = int(f.nnz_in(0)), int(f.nnz_in(1)), int(f.nnz_out(1))
    if int(f.nnz_out(0)) != 1:
        raise ValueError('the first output must be the scalar objective')
    tag = names or 'casadi'
    xsym = [poly.new_symbol('%s_x%d' % (tag, j), 'var') for j in range(n)]
    psym = [poly.new_symbol('%s_p%d' % (tag, k), 'par') for k in range(n_par)]
    xs, ps = [q.single_symbol() for q in xsym], [q.single_symbol() for q in psym]
    inputs = [xsym, psym]
    outputs = [[0.0], [0.0] * m]
    work = [0.0] * int(f.sz_w())
    for k in range(int(f.n_instructions())):
        name = ops.get(f.instruction_id(k))
        o, i = f.instruction_output(k), f.instruction_input(k)
        if name == 'OP_CONST':
            work[o[0]] = float(f.instruction_constant(k))
        elif name == 'OP_INPUT':
            work[o[0]] = inputs[i[0]][i[1]]
        elif name == 'OP_OUTPUT':
            outputs[o[0]][o[1]] = work[i[0]]
        elif name in _BINARY:
            work[o[0]] = _BINARY[name](work[i[0]], work[i[1]])
        elif name in _UNARY:
            work[o[0]] = _UNARY[name](work[i[0]])
        elif name in ('OP_CONSTPOW', 'OP_POW'):
            e = work[i[1]]
            if isinstance(e, Poly) or float(e) != int(float(e)) or float(e) < 0:
                raise NotImplementedError('only non-negative integer powers are polynomial')
            work[o[0]] = _as_poly(work[i[0]]) ** int(float(e))
        elif name == 'OP_IF_ELSE_ZERO':
            work[o[0]] = work[i[0]] * work[i[1]]        # condition is a 0/1 indicator
        elif name == 'OP_NOT':
            work[o[0]] = 1.0 - work[i[0]]
        else:
            raise NotImplementedError('operation %r (code %r) is outside the polynomial + '
                                      'parameter-atom class of the spline NLPs'
                                      % (name, f.instruction_id(k)))
    rows = [_as_poly(r) for r in outputs[1]]
    objective = _as_poly(outputs[0][0])
    lbg = np.broadcast_to(np.asarray(lbg, dtype=float), (m,)).copy()
    ubg = np.broadcast_to(np.asarray(ubg, dtype=float), (m,)).copy()
    return lower(xs, ps, rows, objective, lbg, ubg, order_hint)


def lower_casadi(var, par, obj, con, lbg, ubg, order_hint=None):
    """The reference's ``create_nlp`` arguments (optilayer.py:49-60: var, par, obj, con as MX;
    lbg / ubg = father._lb.cat / father._ub.cat) -> lowered tables."""
    import casadi
    f = casadi.Function('nlp', [var, par], [obj, con]).expand()
    return lower_sx_function(f, lbg, ubg, op_table(casadi), order_hint)
