"""TEST INFRASTRUCTURE: a mechanical executor for the FORTRAN 77 the reference's numeric routines
are written in (lineshape.f: sum_all_lines, humliv_bb, humli_bb; curgods.f: curgod_fort_1..4;
fparts_mod.f: bd_tips_2003 and the QT_* table routines it calls).

No Fortran compiler exists in this image, so the reference's Fortran cannot be compiled.  This
module does the next best thing to running it: it READS THE SOURCE TEXT WHERE IT LIES
(/root/reference/*.f, never copied into the repository), translates each subroutine statement by
statement into a Python function by fixed rules, and executes that.  Nothing about the algorithm is
written here - only the language rules:

  * fixed source form: columns 1-72, comment lines (C, c, *, ! in column 1), `!` comments,
    continuation lines (any character but blank / 0 in column 6), case-insensitive;
  * static types from the declarations (integer*4, real*4, real*8, complex*8, complex*16);
    un-suffixed real literals are REAL*4 (rounded correctly from the decimal text), `d` exponents
    REAL*8; mixed-mode arithmetic promotes integer -> real*4 -> real*8 -> complex as the standard
    says; integer division truncates; assignment converts to the type of the left-hand side;
  * CMPLX(a, b) without a kind is default (single) complex: both parts are rounded to REAL*4;
  * complex*16 multiplication and division as gfortran evaluates them with its default
    -fcx-fortran-rules: the plain four-product form, and division with range reduction
    (Smith: the ratio of the smaller to the larger part of the divisor); no FMA contraction
    (x86-64 baseline);
  * x**n with a literal integer n <= 3 is repeated multiplication (every chain a compiler may
    pick gives the same roundings), with an integer variable n libgfortran's square-and-multiply;
    NINT rounds half away from zero;
  * DO loops evaluate their trip count once and leave the variable one step past the end;
    DO WHILE; block IF / ELSE IF / ELSE; logical IF; STOP raises F77Stop; WRITE is ignored;
  * arrays are passed by reference (NumPy arrays, 1-based subscripts, Fortran order for 2-D);
    a whole-array assignment copies; local arrays with literal bounds are allocated at entry;
  * IMPLICIT statements and the default rule (i-n INTEGER, else REAL*4) - but not in a unit with
    an INCLUDE, whose declarations are not read; DIMENSION; DATA (value lists, repeat counts,
    implied DO over one row of a 2-D array), applied at entry; CALL of another unit of the file,
    scalar arguments copied back by position; GO TO a label that stands on a RETURN.

Everything else - other labels and jumps, COMMON storage not filled by DATA in the same unit,
assignment to a DATA variable (SAVE semantics), array elements as actual arguments, FORMAT, ... -
makes the unit F77Unsupported: it is left out, never guessed.

REAL*8 is a Python float and COMPLEX*16 a Python complex handled component-wise by the helpers
below (IEEE binary64 operations, correctly rounded, the same the compiled code executes).
exp / cos / log go to the C library.

Used by tests/golden/make_f77_golden.py to produce reference-executed fixtures for the rows of
SURVEY section 8 that were pinned by a hand restatement only (A2 humliv_bb, the inner loop of A6
sum_all_lines, the TIPS tables of A9, A13 curgod_fort_*), and by tests/test_f77_golden.py to re-check them whenever
/root/reference is present.
"""
import math
import re
from fractions import Fraction

import numpy as np


class F77Stop(Exception):
    """The translated routine reached a STOP statement."""


class F77Unsupported(Exception):
    """The source uses something outside the subset this executor knows."""


# ---------------------------------------------------------------------------------------------
# run-time helpers the generated code calls
# ---------------------------------------------------------------------------------------------
def _f4(v):
    """Round a REAL*8 value to REAL*4 (kept as the exactly equal Python float)."""
    return float(np.float32(v))


def f4_literal(text):
    """The REAL*4 value of a decimal literal, rounded ONCE from the exact decimal (a compiler
    reads the text at the precision of the kind; going through binary64 could round twice)."""
    exact = Fraction(text)
    c = np.float32(float(exact))
    cands = [c, np.nextafter(c, np.float32(-np.inf)), np.nextafter(c, np.float32(np.inf))]
    # nearest; a tie goes to the even significand
    best = min(cands, key=lambda v: (abs(Fraction(float(v)) - exact), int(v.view(np.uint32)) & 1))
    return float(best)


def _cmul(a, b):
    """COMPLEX*16 product, Fortran rules: (ar*br - ai*bi, ar*bi + ai*br)."""
    return complex(a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real)


def _cmul_r(a, r):
    """COMPLEX*16 times a real: the real operand is a complex with a zero imaginary part."""
    return complex(a.real * r, a.imag * r)


def _cdiv(a, b):
    """COMPLEX*16 quotient with range reduction, as gfortran's -fcx-fortran-rules expands it."""
    if abs(b.real) < abs(b.imag):
        ratio = b.real / b.imag
        div = b.real * ratio + b.imag
        return complex((a.real * ratio + a.imag) / div, (a.imag * ratio - a.real) / div)
    ratio = b.imag / b.real
    div = b.imag * ratio + b.real
    return complex((a.imag * ratio + a.real) / div, (a.imag - a.real * ratio) / div)


def _idiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _ipow(a, n):
    """x**n, integer n >= 1: repeated multiplication from the left (x*x, (x*x)*x, ...)."""
    if n < 1:
        raise F77Unsupported("non-positive integer power")
    r = a
    for _ in range(n - 1):
        r = r * a
    return r


def _ipow_var(a, n):
    """x**n with an integer VARIABLE n: libgfortran's pow_r8_i4 (square and multiply)."""
    if n < 0:
        raise F77Unsupported("negative integer power")
    pw, x, u = (1 if isinstance(a, int) else 1.0), a, n
    while u:
        if u & 1:
            pw = pw * x
        u >>= 1
        if u:
            x = x * x
    return pw


def _nint(v):
    """NINT: nearest integer, halves away from zero."""
    m = abs(v)
    r = math.floor(m)
    if m - r >= 0.5:
        r += 1
    return int(r) if v >= 0 else -int(r)


def _trip(lo, hi, st):
    return max((hi - lo + st) // st, 0) if st > 0 else max((lo - hi - st) // (-st), 0)


def _stop(msg):
    raise F77Stop(msg)


_RUNTIME = dict(np=np, _ipow_var=_ipow_var, _f4=_f4, _cmul=_cmul, _cmul_r=_cmul_r, _cdiv=_cdiv, _idiv=_idiv, _ipow=_ipow,
                _nint=_nint, _trip=_trip, _stop=_stop, math=math, complex=complex, float=float,
                int=int, abs=abs, max=max, min=min, range=range)


# ---------------------------------------------------------------------------------------------
# fixed-form reader
# ---------------------------------------------------------------------------------------------
def _strip_bang(line):
    q = None
    for k, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == '!':
            return line[:k]
    return line


def _lower_outside_quotes(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
            out.append(ch.lower())
    return ''.join(out)


def logical_lines(path):
    """[(first source line number, statement text)] with continuations joined."""
    out = []
    with open(path, errors='replace') as fh:
        for no, raw in enumerate(fh, 1):
            line = raw.rstrip('\r\n').expandtabs(8)[:72]
            if not line.strip() or line[0] in 'Cc*!':
                continue
            line = _strip_bang(line)
            if not line.strip():
                continue
            if len(line) > 5 and line[:5].strip() == '' and line[5] not in ' 0':
                if not out:
                    raise F77Unsupported("continuation without a statement at line %d" % no)
                out[-1] = (out[-1][0], out[-1][1] + line[6:])
                continue
            if line[:5].strip():            # a labelled statement: its unit is outside the subset
                out.append((no, '@label %s %s' % (line[:5].strip(), line[6:].strip())))
                continue
            out.append((no, line[6:]))
    return [(no, _lower_outside_quotes(t).strip()) for no, t in out]


def split_units(path):
    """{subroutine name: (dummy arguments, [(line number, statement)])}"""
    units, cur = {}, None
    for no, st in logical_lines(path):
        m = re.match(r'subroutine\s+(\w+)\s*\((.*)\)\s*$', st)
        if m:
            cur = (m.group(1), [a.strip() for a in m.group(2).split(',')], [])
            continue
        if cur is None:
            continue
        if re.fullmatch(r'end', st):
            units[cur[0]] = (cur[1], cur[2])
            cur = None
            continue
        cur[2].append((no, st))
    return units


# ---------------------------------------------------------------------------------------------
# expressions: tokens -> (python code, type)
# ---------------------------------------------------------------------------------------------
_DOTOPS = ('lt', 'le', 'gt', 'ge', 'eq', 'ne', 'and', 'or', 'not')
_REL = {'lt': '<', 'le': '<=', 'gt': '>', 'ge': '>=', 'eq': '==', 'ne': '!='}
_RANK = {'i': 0, 'r4': 1, 'r8': 2}


def tokenize(s):
    toks, k, n = [], 0, len(s)
    while k < n:
        ch = s[k]
        if ch.isspace():
            k += 1
            continue
        m = re.match(r'\.(%s)\.' % '|'.join(_DOTOPS), s[k:])
        if m:
            toks.append(('op', m.group(1)))
            k += m.end()
            continue
        if ch.isdigit() or (ch == '.' and k + 1 < n and s[k + 1].isdigit()):
            m = re.match(r'\d+', s[k:])
            end = k + (m.end() if m else 0)
            is_real = False
            if end < n and s[end] == '.' and not re.match(r'\.(%s)\.' % '|'.join(_DOTOPS), s[end:]):
                is_real = True
                end += 1
                m2 = re.match(r'\d+', s[end:])
                if m2:
                    end += m2.end()
            mant = s[k:end]
            kind = 'r4' if is_real else 'i'
            m3 = re.match(r'([de])([+-]?\d+)', s[end:])
            expo = ''
            if m3:
                kind = 'r8' if m3.group(1) == 'd' else 'r4'
                expo = 'e' + m3.group(2)
                end += m3.end()
            toks.append(('num', (kind, mant + expo)))
            k = end
            continue
        if ch.isalpha() or ch == '_':
            m = re.match(r'\w+', s[k:])
            toks.append(('id', m.group(0)))
            k += m.end()
            continue
        if s.startswith('**', k):
            toks.append(('op', '**'))
            k += 2
            continue
        if ch in '+-*/(),=':
            toks.append(('op', ch))
            k += 1
            continue
        raise F77Unsupported("character %r in %r" % (ch, s))
    return toks


def _conv(code, frm, to):
    if frm == to:
        return code
    if to == 'r8':
        if frm in ('i', 'r4'):
            return 'float(%s)' % code
    if to == 'r4':
        if frm == 'i':
            return '_f4(%s)' % code
        if frm == 'r8':
            return '_f4(%s)' % code
    if to == 'i' and frm in ('r4', 'r8'):
        return 'int(%s)' % code            # truncation towards zero
    if to == 'c16':
        if frm in ('i', 'r4', 'r8'):
            return 'complex(float(%s), 0.0)' % code
        if frm == 'c8':
            return code                     # parts already REAL*4-valued, widening is exact
    raise F77Unsupported("conversion %s -> %s" % (frm, to))


class ExprParser(object):
    def __init__(self, toks, types, arrays):
        self.t, self.k, self.types, self.arrays = toks, 0, types, arrays

    def peek(self):
        return self.t[self.k] if self.k < len(self.t) else (None, None)

    def take(self, kind=None, val=None):
        tk = self.peek()
        if (kind and tk[0] != kind) or (val is not None and tk[1] != val):
            raise F77Unsupported("expected %s %s, found %s" % (kind, val, tk))
        self.k += 1
        return tk

    def done(self):
        return self.k >= len(self.t)

    # precedence: .or. < .and. < .not. < relational < + - (binary and unary) < * / < **
    def expr(self):
        a = self.and_()
        while self.peek() == ('op', 'or'):
            self.take()
            b = self.and_()
            a = ('(%s or %s)' % (a[0], b[0]), 'l')
        return a

    def and_(self):
        a = self.not_()
        while self.peek() == ('op', 'and'):
            self.take()
            b = self.not_()
            a = ('(%s and %s)' % (a[0], b[0]), 'l')
        return a

    def not_(self):
        if self.peek() == ('op', 'not'):
            self.take()
            a = self.not_()
            return ('(not %s)' % a[0], 'l')
        return self.rel()

    def rel(self):
        a = self.arith()
        tk = self.peek()
        if tk[0] == 'op' and tk[1] in _REL:
            self.take()
            b = self.arith()
            if a[1] not in _RANK or b[1] not in _RANK:
                raise F77Unsupported("comparison of %s and %s" % (a[1], b[1]))
            t = a[1] if _RANK[a[1]] >= _RANK[b[1]] else b[1]
            return ('(%s %s %s)' % (_conv(a[0], a[1], t), _REL[tk[1]], _conv(b[0], b[1], t)), 'l')
        return a

    def arith(self):
        tk = self.peek()
        if tk in (('op', '-'), ('op', '+')):
            self.take()
            a = self.term()
            if tk[1] == '-':
                a = ('(-%s)' % a[0], a[1])
        else:
            a = self.term()
        while self.peek() in (('op', '+'), ('op', '-')):
            op = self.take()[1]
            b = self.term()
            a = self.binary(op, a, b)
        return a

    def term(self):
        a = self.power()
        while self.peek() in (('op', '*'), ('op', '/')):
            op = self.take()[1]
            b = self.power()
            a = self.binary(op, a, b)
        return a

    def power(self):
        a = self.primary()
        if self.peek() == ('op', '**'):
            self.take()
            b = self.power()                      # right-associative
            if b[1] != 'i' or a[1] not in _RANK:
                raise F77Unsupported("non-integer exponent")
            if re.fullmatch(r'\d+', b[0]):
                if not 1 <= int(b[0]) <= 3:      # longer chains: the compiler picks the order
                    raise F77Unsupported("literal integer power above 3")
                return ('_ipow(%s, %s)' % (a[0], b[0]), a[1])
            return ('_ipow_var(%s, %s)' % (a[0], b[0]), a[1])
        return a

    @staticmethod
    def result_type(ta, tb):
        cplx = [t for t in (ta, tb) if t.startswith('c')]
        if cplx:
            if 'c16' in cplx or 'r8' in (ta, tb):
                return 'c16'
            return 'c8'
        return ta if _RANK[ta] >= _RANK[tb] else tb

    def binary(self, op, a, b):
        t = self.result_type(a[1], b[1])
        if t == 'c8':
            raise F77Unsupported("single-precision complex arithmetic")
        if t == 'c16':
            if op in '+-':
                # complex +- real touches the real part only; complex +- complex both
                ca = a[0] if a[1].startswith('c') else _conv(a[0], a[1], 'c16')
                cb = b[0] if b[1].startswith('c') else _conv(b[0], b[1], 'c16')
                return ('(%s %s %s)' % (ca, op, cb), 'c16')
            if op == '*':
                if a[1].startswith('c') and b[1].startswith('c'):
                    return ('_cmul(%s, %s)' % (a[0], b[0]), 'c16')
                c, r = (a, b) if a[1].startswith('c') else (b, a)
                return ('_cmul_r(%s, %s)' % (c[0], _conv(r[0], r[1], 'r8')), 'c16')
            if a[1].startswith('c') and b[1].startswith('c'):
                return ('_cdiv(%s, %s)' % (a[0], b[0]), 'c16')
            raise F77Unsupported("complex / real or real / complex")
        ca, cb = _conv(a[0], a[1], t), _conv(b[0], b[1], t)
        if t == 'i':
            return ('_idiv(%s, %s)' % (ca, cb), 'i') if op == '/' else ('(%s %s %s)' % (ca, op, cb), 'i')
        code = '(%s %s %s)' % (ca, op, cb)
        return ('_f4%s' % code if t == 'r4' else code, t)

    def args(self):
        self.take('op', '(')
        out = []
        if self.peek() != ('op', ')'):
            out.append(self.expr())
            while self.peek() == ('op', ','):
                self.take()
                out.append(self.expr())
        self.take('op', ')')
        return out

    def primary(self):
        kind, val = self.peek()
        if kind == 'num':
            self.take()
            k, text = val
            if k == 'i':
                return (text.lstrip('0') or '0', 'i')
            if k == 'r8':
                return (repr(float(text)), 'r8')
            return (repr(f4_literal(text)), 'r4')
        if kind == 'op' and val == '(':
            self.take()
            a = self.expr()
            self.take('op', ')')
            return ('(%s)' % a[0], a[1])
        if kind == 'id':
            self.take()
            if self.peek() == ('op', '('):
                a = self.args()
                if val in self.arrays:
                    if any(x[1] != 'i' for x in a) or len(a) != self.arrays[val]:
                        raise F77Unsupported("subscripts of %s" % val)
                    self.types.used.add(val)
                    return ('%s[%s]' % (val, ', '.join('%s - 1' % x[0] for x in a)), self.types[val])
                return self.intrinsic(val, a)
            if self.types.lookup(val) is None:
                raise F77Unsupported("undeclared name %s" % val)
            self.types.used.add(val)
            if val in self.arrays:
                return (val, 'array:' + self.types[val])
            return (val, self.types[val])
        raise F77Unsupported("unexpected token %s %s" % (kind, val))

    def intrinsic(self, name, a):
        t0 = a[0][1]
        if name == 'cmplx' and len(a) == 2:
            return ('complex(_f4(%s), _f4(%s))' % (_conv(a[0][0], a[0][1], 'r8'),
                                                     _conv(a[1][0], a[1][1], 'r8')), 'c8')
        if name == 'dble' and len(a) == 1:
            return ('(%s).real' % a[0][0], 'r8') if t0.startswith('c') else (_conv(a[0][0], t0, 'r8'), 'r8')
        if name == 'dimag' and len(a) == 1 and t0 == 'c16':
            return ('(%s).imag' % a[0][0], 'r8')
        if name in ('exp', 'cos', 'sin', 'log', 'sqrt') and len(a) == 1 and t0 in ('r4', 'r8'):
            code = 'math.%s(%s)' % (name, a[0][0])
            return ('_f4(%s)' % code if t0 == 'r4' else code, t0)
        if name == 'abs' and len(a) == 1 and t0 in _RANK:
            return ('abs(%s)' % a[0][0], t0)
        if name == 'nint' and len(a) == 1 and t0 in ('r4', 'r8'):
            return ('_nint(%s)' % a[0][0], 'i')
        if name in ('max', 'min') and len(a) >= 2:
            t = a[0][1]
            for x in a[1:]:
                t = self.result_type(t, x[1])
            if t not in _RANK:
                raise F77Unsupported("max/min of %s" % t)
            return ('%s(%s)' % (name, ', '.join(_conv(x[0], x[1], t) for x in a)), t)
        raise F77Unsupported("intrinsic %s/%d" % (name, len(a)))


# ---------------------------------------------------------------------------------------------
# statements -> python source
# ---------------------------------------------------------------------------------------------
_TYPES = [(r'real\s*\*\s*8|double\s+precision', 'r8'), (r'real\s*\*\s*4|real', 'r4'),
          (r'integer\s*\*\s*4|integer', 'i'), (r'complex\s*\*\s*16|double\s+complex', 'c16'),
          (r'complex\s*\*\s*8|complex', 'c8')]


def _split_top(s):
    out, depth, cur = [], 0, ''
    for ch in s:
        if ch == '(':
            depth += 1
        elif ch == ')':
            depth -= 1
        if ch == ',' and depth == 0:
            out.append(cur)
            cur = ''
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def _matching_paren(s, k):
    depth = 0
    for j in range(k, len(s)):
        if s[j] == '(':
            depth += 1
        elif s[j] == ')':
            depth -= 1
            if depth == 0:
                return j
    raise F77Unsupported("unbalanced parentheses in %r" % s)


class TypeTable(dict):
    """name -> type, with the IMPLICIT rule of the unit for names that were never declared
    (`implicit none`, or an INCLUDE whose declarations are not read: undeclared names are
    refused)."""

    def __init__(self):
        dict.__init__(self)
        # the standard's default: i-n INTEGER, everything else REAL
        self.implicit = dict((chr(c), 'i' if chr(c) in 'ijklmn' else 'r4')
                             for c in range(ord('a'), ord('z') + 1))
        self.used = set()

    def lookup(self, name):
        if name in self:
            return self[name]
        if self.implicit is None:
            return None
        self[name] = self.implicit[name[0]]
        return self[name]


class Translator(object):
    def __init__(self, name, dummies, stmts):
        self.name, self.dummies, self.stmts = name, dummies, stmts
        self.types, self.arrays = TypeTable(), {}
        self.dims, self.common, self.data = {}, set(), []
        self.lines, self.depth, self.tmp = [], 1, 0
        self.return_labels = set()

    def emit(self, code):
        self.lines.append('    ' * self.depth + code)

    def parse(self, text):
        p = ExprParser(tokenize(text), self.types, self.arrays)
        r = p.expr()
        if not p.done():
            raise F77Unsupported("trailing tokens in %r" % text)
        return r

    def declare_items(self, text, t):
        for item in _split_top(text):
            mm = re.match(r'(\w+)\s*(\((.*)\))?$', item)
            if not mm:
                raise F77Unsupported("declaration %r" % text)
            nam = mm.group(1)
            if t is not None:
                self.types[nam] = t
            elif self.types.lookup(nam) is None:
                raise F77Unsupported("dimension of the untyped name %s" % nam)
            if mm.group(2):
                dims = _split_top(mm.group(3))
                self.arrays[nam] = len(dims)
                self.dims[nam] = [d.strip() for d in dims]

    def declaration(self, st):
        if st == 'implicit none':
            self.types.implicit = None
            return True
        m = re.match(r'implicit\s+(.*?)\s*\(([a-z,\s-]+)\)\s*$', st)
        if m:
            t = [tt for pat, tt in _TYPES if re.fullmatch(pat, m.group(1))]
            if not t:
                raise F77Unsupported("statement %r" % st)
            rule = dict((chr(c), 'i' if chr(c) in 'ijklmn' else 'r4') for c in range(ord('a'), ord('z') + 1))
            for rng in m.group(2).split(','):
                lo, _, hi = rng.strip().partition('-')
                for c in range(ord(lo.strip()), ord((hi or lo).strip()) + 1):
                    rule[chr(c)] = t[0]
            self.types.implicit = rule
            return True
        if st.startswith('include'):               # its declarations are not read, so an undeclared
            self.types.implicit = None             # name may be one of them: refuse, do not guess
            return True
        m = re.match(r'common\s*/\s*\w*\s*/(.*)$', st)
        if m:                                      # storage association is not modelled: a COMMON
            for nam in _split_top(m.group(1)):     # variable may only be used where DATA fills it
                self.common.add(re.match(r'\w+', nam).group(0))
            return True
        m = re.match(r'dimension\s+(.*)$', st)
        if m:
            self.declare_items(m.group(1), None)
            return True
        m = re.match(r'data\s*(.*)$', st)
        if m and '/' in st:
            self.data.append(m.group(1))
            return True
        for pat, t in _TYPES:
            m = re.match(r'(?:%s)\s+(?![=(])' % pat, st)
            if m:
                self.declare_items(st[m.end():], t)
                return True
        return False

    def data_values(self, text, t):
        vals = []
        for item in _split_top(text):
            rep, _, v = item.rpartition('*')
            toks = tokenize(v)
            sign = 1
            if toks and toks[0] in (('op', '-'), ('op', '+')):
                sign = -1 if toks[0][1] == '-' else 1
                toks = toks[1:]
            if len(toks) != 1 or toks[0][0] != 'num':
                raise F77Unsupported("DATA value %r" % item)
            kind, txt = toks[0][1]
            x = int(txt) if kind == 'i' else (float(txt) if kind == 'r8' else f4_literal(txt))
            x = sign * x
            if t == 'i':
                if kind != 'i':
                    raise F77Unsupported("real DATA value for an integer")
            elif t == 'r4':
                x = _f4(x)
            elif t == 'r8':
                x = float(x)
            else:
                raise F77Unsupported("DATA for type %s" % t)
            vals += [x] * (int(rep) if rep.strip() else 1)
        return vals

    def emit_data(self):
        """DATA statements: `name / list /` (a scalar or a whole array in storage order) and
        `(name(k, j), j = a, b) / list /`.  Emitted as initialisation at entry; a DATA variable
        that the body assigns would need SAVE semantics and is refused (see assignment)."""
        done = set()
        for text in self.data:
            m = re.match(r'\(\s*(\w+)\s*\(\s*(\d+)\s*,\s*(\w+)\s*\)\s*,\s*(\w+)\s*=\s*(\d+)\s*,\s*(\d+)\s*\)\s*/(.*)/\s*$', text)
            if m and m.group(3) == m.group(4) and self.arrays.get(m.group(1)) == 2:
                nam, row, lo, hi = m.group(1), int(m.group(2)), int(m.group(5)), int(m.group(6))
                vals = self.data_values(m.group(7), self.types[nam])
                if len(vals) != hi - lo + 1:
                    raise F77Unsupported("DATA count for %s" % nam)
                self.emit('%s[%d, %d:%d] = %r' % (nam, row - 1, lo - 1, hi, vals))
                done.add(nam)
                continue
            m = re.match(r'(\w+)\s*/(.*)/\s*$', text)
            if not m or self.types.lookup(m.group(1)) is None:
                raise F77Unsupported("DATA statement %r" % text)
            nam = m.group(1)
            vals = self.data_values(m.group(2), self.types[nam])
            if nam in self.arrays:
                if self.arrays[nam] != 1 or len(vals) != int(self.dims[nam][0]):
                    raise F77Unsupported("DATA count for %s" % nam)
                self.emit('%s[:] = %r' % (nam, vals))
            else:
                if len(vals) != 1:
                    raise F77Unsupported("DATA count for %s" % nam)
                self.emit('%s = %r' % (nam, vals[0]))
            done.add(nam)
        return done

    def assignment(self, st):
        depth = 0
        for k, ch in enumerate(st):
            depth += (ch == '(') - (ch == ')')
            if ch == '=' and depth == 0:
                break
        else:
            raise F77Unsupported("statement %r" % st)
        lhs, rhs = st[:k].strip(), st[k + 1:].strip()
        m = re.match(r'(\w+)\s*(\(.*\))?$', lhs)
        if not m or self.types.lookup(m.group(1)) is None:
            raise F77Unsupported("left-hand side %r" % lhs)
        nam, t = m.group(1), self.types[m.group(1)]
        if nam in self.data_names:
            raise F77Unsupported("assignment to the DATA variable %s" % nam)
        val = self.parse(rhs)
        if nam in self.arrays and not m.group(2):              # whole-array assignment
            if val[1] != 'array:' + t:
                raise F77Unsupported("array assignment %r" % st)
            self.emit('%s[...] = %s' % (nam, val[0]))
            return
        code = _conv(val[0], val[1], t)
        if m.group(2):
            target = self.parse(lhs)[0]
        else:
            target = nam
        self.emit('%s = %s' % (target, code))

    def call(self, st):
        """CALL name(a, b, ...): actual arguments must be plain names; arrays travel by reference,
        scalars are copied back from the callee's final values (by position)."""
        m = re.match(r'call\s+(\w+)\s*\((.*)\)\s*$', st)
        if not m:
            raise F77Unsupported("statement %r" % st)
        args = _split_top(m.group(2))
        for a in args:
            if not re.fullmatch(r'\w+', a) or self.types.lookup(a) is None:
                raise F77Unsupported("actual argument %r" % a)
            self.types.used.add(a)
        self.tmp += 1
        r = '_call%d' % self.tmp
        self.emit('%s = _call(_units, %r, [%s])' % (r, m.group(1), ', '.join(args)))
        for k, a in enumerate(args):
            if a not in self.arrays:
                if a in self.data_names:
                    raise F77Unsupported("DATA variable %s as an actual argument" % a)
                self.emit('%s = %s[%d]' % (a, r, k))

    def simple(self, st):
        if st.startswith('stop'):
            self.emit('_stop(%r)' % st[4:].strip().strip("'\""))
        elif re.match(r'write\s*\(', st) or re.match(r'print\b', st):
            self.emit('pass')
        elif st == 'return':
            self.emit('return _result()')
        elif re.match(r'call\s', st):
            self.call(st)
        elif re.fullmatch(r'go\s*to\s*(\d+)', st):
            lab = re.fullmatch(r'go\s*to\s*(\d+)', st).group(1)
            if lab not in self.return_labels:      # only a jump to a labelled RETURN is known
                raise F77Unsupported("go to %s" % lab)
            self.emit('return _result()')
        else:
            self.assignment(st)

    def translate(self):
        body = []
        for no, st in self.stmts:
            m = re.match(r'@label (\d+) (.*)$', st)
            if m:
                if m.group(2).strip() != 'return':
                    raise F77Unsupported("statement label at line %d" % no)
                self.return_labels.add(m.group(1))
                body.append((no, 'return'))
                continue
            if not body and self.declaration(st):
                continue
            if body and (re.match(r'(implicit|dimension|common|data)\b', st) or
                         any(re.match(r'(?:%s)\s+(?![=(])' % pat, st) for pat, _ in _TYPES)):
                if not self.declaration(st):
                    raise F77Unsupported("statement %r" % st)
                continue
            body.append((no, st))
        for d in self.dummies:
            if self.types.lookup(d) is None:
                raise F77Unsupported("dummy %s undeclared" % d)
        self.lines.append('def %s(%s):' % (self.name, ', '.join(self.dummies)))
        self.data_names = set()
        for text in self.data:                    # names first: the body must not assign them
            m = re.match(r'\(\s*(\w+)', text) or re.match(r'(\w+)', text)
            self.data_names.add(m.group(1))
        if self.data_names & set(self.dummies):
            raise F77Unsupported("DATA for a dummy argument")
        body_lines, self.lines = self.lines, []
        self.body(body)
        code, self.lines = self.lines, body_lines
        local = [n for n in self.types if n not in self.dummies]
        for n in local:
            if n in self.arrays:
                try:
                    shape = tuple(int(d) for d in self.dims[n])
                except ValueError:
                    if n in self.types.used:
                        raise F77Unsupported("local array %s with symbolic bounds" % n)
                    continue
                dt = {'i': 'np.int64', 'r4': 'np.float64', 'r8': 'np.float64'}.get(self.types[n])
                if dt is None:
                    raise F77Unsupported("local array %s of type %s" % (n, self.types[n]))
                self.emit("%s = np.zeros(%r, dtype=%s, order='F')" % (n, shape, dt))
            else:
                self.emit('%s = None' % n)
        inited = self.emit_data()
        for n in self.common:
            if n in self.types.used and n not in inited:
                raise F77Unsupported("COMMON variable %s used without DATA in this unit" % n)
        scal = [d for d in self.dummies if d not in self.arrays]
        self.emit('def _result():')
        self.emit('    return {%s}' % ', '.join('%r: %s' % (d, d) for d in scal))
        for d in scal:               # scalar dummies arrive as Python numbers of their declared type
            if self.types[d] not in ('i', 'r8'):
                raise F77Unsupported("scalar dummy %s of type %s" % (d, self.types[d]))
            self.emit('%s = %s(%s)' % (d, 'int' if self.types[d] == 'i' else 'float', d))
        self.lines += code
        self.emit('return _result()')
        return '\n'.join(self.lines) + '\n'

    def body(self, body):
        stack = []
        for no, st in body:
            self.emit('# line %d' % no)
            m = re.match(r'(else\s*if|if)\s*\(', st)
            if m:
                close = _matching_paren(st, m.end() - 1)
                cond = self.parse(st[m.end():close])
                if cond[1] != 'l':
                    raise F77Unsupported("condition %r" % st)
                rest = st[close + 1:].strip()
                if m.group(1) == 'if':
                    if rest == 'then':
                        self.emit('if %s:' % cond[0])
                        stack.append('if')
                        self.depth += 1
                        self.emit('pass')
                    else:
                        self.emit('if %s:' % cond[0])
                        self.depth += 1
                        self.simple(rest)
                        self.depth -= 1
                else:
                    if rest != 'then' or not stack or stack[-1] != 'if':
                        raise F77Unsupported("else if at line %d" % no)
                    self.depth -= 1
                    self.emit('elif %s:' % cond[0])
                    self.depth += 1
                    self.emit('pass')
                continue
            if st == 'else':
                if not stack or stack[-1] != 'if':
                    raise F77Unsupported("else at line %d" % no)
                self.depth -= 1
                self.emit('else:')
                self.depth += 1
                self.emit('pass')
                continue
            if re.fullmatch(r'end\s*if', st):
                if not stack or stack.pop() != 'if':
                    raise F77Unsupported("end if at line %d" % no)
                self.depth -= 1
                continue
            m = re.match(r'do\s+while\s*\(', st)
            if m:
                close = _matching_paren(st, m.end() - 1)
                cond = self.parse(st[m.end():close])
                self.emit('while %s:' % cond[0])
                stack.append(('while', None))
                self.depth += 1
                self.emit('pass')
                continue
            m = re.match(r'do\s+(\w+)\s*=(.*)$', st)
            if m:
                var = m.group(1)
                parts = _split_top(m.group(2))
                if self.types.lookup(var) != 'i' or len(parts) not in (2, 3):
                    raise F77Unsupported("do statement %r" % st)
                ex = [self.parse(p) for p in parts]
                if any(e[1] != 'i' for e in ex):
                    raise F77Unsupported("non-integer do bounds %r" % st)
                self.tmp += 1
                t = '_do%d' % self.tmp
                step = ex[2][0] if len(ex) == 3 else '1'
                self.emit('%s_st = %s' % (t, step))
                self.emit('%s = %s' % (var, ex[0][0]))
                self.emit('for %s_k in range(_trip(%s, %s, %s_st)):' % (t, var, ex[1][0], t))
                stack.append(('do', (var, t)))
                self.depth += 1
                self.emit('pass')
                continue
            if re.fullmatch(r'end\s*do', st):
                if not stack or stack[-1] == 'if':
                    raise F77Unsupported("end do at line %d" % no)
                kind, info = stack.pop()
                if kind == 'do':
                    self.emit('%s = %s + %s_st' % (info[0], info[0], info[1]))
                self.depth -= 1
                continue
            self.simple(st)
        if stack:
            raise F77Unsupported("unterminated block in %s" % self.name)


def _call(units, name, actuals):
    """Run-time side of CALL: the callee's final scalar values by position (None for arrays)."""
    fn = units.get(name)
    if fn is None:
        raise F77Unsupported("call of %s, which is outside the subset" % name)
    if len(actuals) != len(fn.dummies):
        raise F77Unsupported("argument count in the call of %s" % name)
    res = fn(*actuals)
    return [res.get(d) for d in fn.dummies]


_CACHE = {}


def load(path):
    """{name: python function} for every subroutine of the file the executor can translate;
    `.source` on each function holds the generated Python, `.dummies` the argument names."""
    if path in _CACHE:
        return _CACHE[path]
    out = {}
    for name, (dummies, stmts) in split_units(path).items():
        try:
            src = Translator(name, dummies, stmts).translate()
        except F77Unsupported:
            continue
        ns = dict(_RUNTIME)
        ns['_units'] = out
        ns['_call'] = _call
        exec(compile(src, '<f77:%s:%s>' % (path, name), 'exec'), ns)
        fn = ns[name]
        fn.source = src
        fn.dummies = list(dummies)
        out[name] = fn
    _CACHE[path] = out
    return out
