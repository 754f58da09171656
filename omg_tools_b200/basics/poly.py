"""Sparse multivariate polynomials with float coefficients -- the symbolic
scalar type of this framework.

Role: the reference models every NLP with CasADi ``MX`` symbols
(omgtools/basics/optilayer.py:556-629) and lets CasADi differentiate the
resulting graph.  CasADi is not part of this framework: every spline NLP of the
hot path is a *polynomial* in the decision variables whose coefficients are
functions of the parameters only, so the symbolic scalar here is a canonical
sparse polynomial over a global symbol registry.  Symbols are

* ``var``  -- decision variables (spline coefficients, slacks, hyperplanes)
* ``par``  -- parameters (T, t, state0, obstacle x/v/a, ...)
* ``sym``  -- named placeholders resolved by name at problem composition
              (reference: OptiChild.define_symbol / OptiFather.translate_symbols,
              optilayer.py:204-223, 556-557)
* ``atom`` -- a non-polynomial function of parameter-only polynomials
              (1/p, indicator p>=0 / p>0, sin, cos, sqrt).  They appear in
              ``t/T`` (point2point.py:55), the Cox-de Boor indicators of
              ``evalspline`` (spline_extra.py:28-55) and rotating obstacles
              (obstacle.py:292-332).  The identity atom ``id`` names a
              parameter-only polynomial so that it is evaluated once (what a
              shared node of CasADi's expression graph is).
* ``mid``  -- a named intermediate: a polynomial of the decision variables
              that many rows share (the product-spline coefficients of
              Quadrotor3D's accelerations, quadrotor3d.py:91-93).  Rows must
              be affine in mids with parameter-only coefficients; lowering.py
              differentiates through them by the chain rule instead of
              expanding them into every row.

A ``Poly`` is a dict {monomial: coef}; a monomial is a sorted tuple of symbol
ids (repetition = power).  ``lowering.py`` turns rows of Poly into the flat
tables the CUDA kernels consume.
"""
import math
import numbers

import numpy as np

# --------------------------------------------------------------------------
# symbol registry
# --------------------------------------------------------------------------


class SymInfo(object):
    __slots__ = ('id', 'name', 'kind', 'func', 'arg', 'alias')

    def __init__(self, id_, name, kind, func=None, arg=None):
        self.id = id_
        self.name = name
        self.kind = kind
        self.func = func
        self.arg = arg
        self.alias = None

    def __repr__(self):
        return '<%s %s#%d>' % (self.kind, self.name, self.id)


_SYMS = []
_ATOMS = {}

ATOM_FUNCS = ('id', 'inv', 'ge', 'gt', 'sin', 'cos', 'sqrt')


def sym_info(sid):
    return _SYMS[sid]


def resolve(sid):
    """Follow placeholder aliases to the defining symbol id."""
    info = _SYMS[sid]
    while info.alias is not None:
        info = _SYMS[info.alias]
    return info.id


def new_symbol(name, kind):
    if kind not in ('var', 'par', 'sym'):
        raise ValueError('unknown symbol kind %s' % kind)
    info = SymInfo(len(_SYMS), name, kind)
    _SYMS.append(info)
    return Poly({(info.id,): 1.0})


def new_mid(name, definition):
    """Intermediate symbol standing for the Poly ``definition``."""
    info = SymInfo(len(_SYMS), name, 'mid', None, definition)
    _SYMS.append(info)
    return Poly({(info.id,): 1.0})


def rel_time(t, T):
    """t / T, the relative start of the horizon.  When the motion time T is a decision
    VARIABLE (FreeTPoint2point) the reference's parameter t is identically 0 -- the time axis
    resets at every update, point2point.py:300-306 -- and t / T is not polynomial in T:
    the relative time is then the constant 0."""
    if isinstance(T, Poly) and len(T.t) == 1:
        (mono, c), = T.t.items()
        if len(mono) == 1 and c == 1.0 and _SYMS[resolve(mono[0])].kind == 'var':
            return 0.
    return t / T


def share(poly):
    """Parameter-only Poly -> one symbol (identity atom) evaluated once."""
    poly = _as_poly(poly)
    if poly.is_constant() or (len(poly.t) == 1 and abs(list(poly.t.values())[0] - 1.) == 0.
                              and len(list(poly.t)[0]) == 1):
        return poly
    return _atom('id', poly)


def collapse(poly):
    """Rewrite a Poly that is affine in its var/mid symbols so that every
    var/mid monomial carries ONE parameter symbol as coefficient (its former
    parameter polynomial, shared through an identity atom).  Placeholder
    symbols count as parameters here (fixed-T problems)."""
    if not isinstance(poly, Poly):
        return poly
    groups = {}
    for mono, c in poly.t.items():
        xm = tuple(s for s in mono if _SYMS[resolve(s)].kind in ('var', 'mid'))
        pm = tuple(s for s in mono if _SYMS[resolve(s)].kind not in ('var', 'mid'))
        groups.setdefault(xm, {})[pm] = c
    out = Poly()
    for xm, pp in groups.items():
        coef = Poly(pp)
        if len(pp) > 1 and xm:
            coef = share(coef)
        out = out + coef * Poly({xm: 1.0})
    return out


def set_alias(placeholder, target):
    """Resolve placeholder symbol (a Poly holding one 'sym') to target."""
    pid = placeholder.single_symbol()
    tid = target.single_symbol()
    if _SYMS[pid].kind != 'sym':
        raise ValueError('%r is not a placeholder' % _SYMS[pid])
    _SYMS[pid].alias = tid


def _atom(func, arg):
    key = (func, arg.key())
    sid = _ATOMS.get(key)
    if sid is None:
        info = SymInfo(len(_SYMS), func, 'atom', func, arg)
        _SYMS.append(info)
        sid = info.id
        _ATOMS[key] = sid
    return Poly({(sid,): 1.0})


# --------------------------------------------------------------------------
# polynomial
# --------------------------------------------------------------------------

def _elementwise(fun, arr):
    out = np.empty(arr.shape, dtype=object)
    of, af = out.reshape(-1), arr.reshape(-1)
    for k in range(af.size):
        of[k] = fun(af[k])
    return out


def _is_array(other):
    return isinstance(other, np.ndarray) and other.size != 1


def _as_poly(other):
    if isinstance(other, Poly):
        return other
    if isinstance(other, numbers.Real):
        c = float(other)
        return Poly({(): c} if c != 0.0 else {})
    if isinstance(other, np.ndarray) and other.size == 1:
        return _as_poly(other.reshape(-1)[0])
    return None


class Poly(object):
    __slots__ = ('t',)
    # make numpy scalars / arrays defer to our reflected operators
    __array_ufunc__ = None

    def __init__(self, terms=None):
        self.t = terms if terms is not None else {}

    # -- inspection -------------------------------------------------------
    def key(self):
        return tuple(sorted(self.t.items()))

    def is_constant(self):
        return all(len(m) == 0 for m in self.t)

    def constant_value(self):
        return self.t.get((), 0.0)

    def single_symbol(self):
        if len(self.t) != 1:
            raise ValueError('not a single symbol: %r' % self)
        (mono, c), = self.t.items()
        if len(mono) != 1 or c != 1.0:
            raise ValueError('not a single symbol: %r' % self)
        return mono[0]

    def symbols(self):
        out = set()
        for mono in self.t:
            out.update(mono)
        return out

    def degree(self):
        return max([len(m) for m in self.t] + [0])

    def __repr__(self):
        if not self.t:
            return 'Poly(0)'
        parts = []
        for mono, c in sorted(self.t.items()):
            names = '*'.join('%s#%d' % (_SYMS[s].name, s) for s in mono)
            parts.append('%+.6g%s' % (c, ('*' + names) if names else ''))
        return 'Poly(' + ' '.join(parts) + ')'

    def __hash__(self):
        return hash(self.key())

    def __eq__(self, other):
        other = _as_poly(other)
        if other is None:
            return NotImplemented
        return self.t == other.t

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __bool__(self):
        return bool(self.t)

    # -- ring operations --------------------------------------------------
    def __neg__(self):
        return Poly({m: -c for m, c in self.t.items()})

    def __pos__(self):
        return self

    def __add__(self, other):
        if _is_array(other):
            return _elementwise(lambda e: self + e, other)
        other = _as_poly(other)
        if other is None:
            return NotImplemented
        if not other.t:
            return self
        if not self.t:
            return other
        res = dict(self.t)
        for m, c in other.t.items():
            v = res.get(m, 0.0) + c
            if v == 0.0:
                res.pop(m, None)
            else:
                res[m] = v
        return Poly(res)

    __radd__ = __add__

    def __sub__(self, other):
        if _is_array(other):
            return _elementwise(lambda e: self - e, other)
        other = _as_poly(other)
        if other is None:
            return NotImplemented
        return self + (-other)

    def __rsub__(self, other):
        if _is_array(other):
            return _elementwise(lambda e: e - self, other)
        other = _as_poly(other)
        if other is None:
            return NotImplemented
        return other + (-self)

    def __mul__(self, other):
        if isinstance(other, numbers.Real):
            c = float(other)
            if c == 0.0:
                return Poly()
            if c == 1.0:
                return self
            return Poly({m: v * c for m, v in self.t.items()})
        if _is_array(other):
            return _elementwise(lambda e: self * e, other)
        other = _as_poly(other)
        if other is None:
            return NotImplemented
        if not self.t or not other.t:
            return Poly()
        res = {}
        for m1, c1 in self.t.items():
            for m2, c2 in other.t.items():
                if not m1:
                    m = m2
                elif not m2:
                    m = m1
                else:
                    m = tuple(sorted(m1 + m2))
                v = res.get(m, 0.0) + c1 * c2
                if v == 0.0:
                    res.pop(m, None)
                else:
                    res[m] = v
        return Poly(res)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, numbers.Real):
            return self * (1.0 / float(other))
        if _is_array(other):
            return _elementwise(lambda e: self / e, other)
        other = _as_poly(other)
        if other is None:
            return NotImplemented
        if other.is_constant():
            return self * (1.0 / other.constant_value())
        return self * _atom('inv', other)

    def __rtruediv__(self, other):
        if _is_array(other):
            return _elementwise(lambda e: e / self, other)
        other = _as_poly(other)
        if other is None:
            return NotImplemented
        return other / self

    def __pow__(self, power):
        if not isinstance(power, numbers.Integral) or power < 0:
            raise TypeError('Poly power must be a non-negative integer')
        res = Poly({(): 1.0})
        for _ in range(int(power)):
            res = res * self
        return res

    # -- comparisons give 0/1 indicator atoms ------------------------------
    # (used by the Cox-de Boor recursion with symbolic abscissa,
    #  reference spline_extra.py:37-41)
    def __ge__(self, other):
        return _indicator('ge', self - other)

    def __gt__(self, other):
        return _indicator('gt', self - other)

    def __le__(self, other):
        return _indicator('ge', other - self)

    def __lt__(self, other):
        return _indicator('gt', other - self)

    # -- numeric evaluation (host-side checks only) -------------------------
    def evaluate(self, values):
        """values: dict {resolved symbol id: float}; atoms evaluated lazily."""
        total = 0.0
        for mono, c in self.t.items():
            prod = c
            for s in mono:
                prod *= _value_of(s, values)
            total += prod
        return total


def _value_of(sid, values):
    sid = resolve(sid)
    if sid in values:
        return values[sid]
    info = _SYMS[sid]
    if info.kind == 'mid':
        v = info.arg.evaluate(values)
        values[sid] = v
        return v
    if info.kind != 'atom':
        raise KeyError('no value for %r' % info)
    v = apply_atom(info.func, info.arg.evaluate(values))
    values[sid] = v
    return v


def apply_atom(func, a):
    if func == 'id':
        return a
    if func == 'inv':
        return 1.0 / a
    if func == 'ge':
        return 1.0 if a >= 0.0 else 0.0
    if func == 'gt':
        return 1.0 if a > 0.0 else 0.0
    if func == 'sin':
        return math.sin(a)
    if func == 'cos':
        return math.cos(a)
    if func == 'sqrt':
        return math.sqrt(a)
    raise ValueError(func)


def _indicator(func, arg):
    arg = _as_poly(arg)
    if arg.is_constant():
        return _as_poly(apply_atom(func, arg.constant_value()))
    return _atom(func, arg)


def substitute(expr, mapping):
    """Replace symbols (ids in ``mapping``) of a Poly by Poly expressions."""
    if not isinstance(expr, Poly):
        return expr
    if not any(s in mapping for s in expr.symbols()):
        return expr
    res = Poly()
    for mono, c in expr.t.items():
        term = Poly({(): c})
        rest = []
        for s in mono:
            if s in mapping:
                term = term * mapping[s]
            else:
                rest.append(s)
        if rest:
            term = term * Poly({tuple(rest): 1.0})
        res = res + term
    return res


def _unary(func, x):
    if isinstance(x, numbers.Real):
        return apply_atom(func, float(x))
    p = _as_poly(x)
    if p is None:
        raise TypeError('cannot apply %s to %r' % (func, x))
    if p.is_constant():
        return apply_atom(func, p.constant_value())
    return _atom(func, p)


def sin(x):
    return _unary('sin', x)


def cos(x):
    return _unary('cos', x)


def sqrt(x):
    return _unary('sqrt', x)


# --------------------------------------------------------------------------
# arrays of Poly (the MX-matrix stand-in)
# --------------------------------------------------------------------------

def sym_array(name, kind, size0, size1=1):
    """Column-major named block of fresh symbols, like ``MX.sym(name,n,m)``.

    Returned as an object ndarray of shape (size0, size1); element (i, j) is
    flat entry j*size0+i of the block (casadi column-major flattening,
    cf. reference Point2Point.cpp:250,290).
    """
    arr = np.empty((size0, size1), dtype=object)
    for j in range(size1):
        for i in range(size0):
            arr[i, j] = new_symbol('%s[%d,%d]' % (name, i, j), kind)
    return arr


def is_symbolic(x):
    if isinstance(x, Poly):
        return True
    if isinstance(x, np.ndarray) and x.dtype == object:
        return True
    return False


def matvec(T, v):
    """T (dense ndarray or scipy sparse) times vector v (float or Poly)."""
    if hasattr(T, 'toarray'):
        T = T.toarray()
    T = np.asarray(T, dtype=float)
    v = np.asarray(v)
    if v.dtype != object:
        return T.dot(v.astype(float))
    out = np.empty(T.shape[0], dtype=object)
    for i in range(T.shape[0]):
        acc = Poly()
        row = T[i]
        for j in np.nonzero(row)[0]:
            acc = acc + v[j] * row[j]
        out[i] = acc
    return out


def as_object_vector(v):
    v = np.asarray(v)
    if v.dtype == object:
        return v
    out = np.empty(v.shape, dtype=object)
    flat = out.reshape(-1)
    for k, x in enumerate(v.reshape(-1)):
        flat[k] = _as_poly(float(x))
    return out
