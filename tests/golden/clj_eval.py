"""A minimal evaluator for the Clojure subset that jvia/amclj's particle-filter hot path is written in.

TEST INFRASTRUCTURE.  Purpose: pin the CPU oracle (oracle/amclj_oracle.c) to the REFERENCE'S OWN SOURCE
TEXT.  No JVM exists in this image, so the reference cannot run; but its hot path uses a small part of the
language — defn / fn / #(), let with destructuring, loop / recur, if / cond / when, for, ->, get-in /
update-in / assoc, keywords as functions, a handful of sequence functions and java.lang.Math — and that
part is executed here directly from `/root/reference/src/amclj/{quaternions,map,motion,sensor,pf}.clj`,
read verbatim: nothing of the reference is retyped.  What is bound from the outside is exactly what the
hot path calls out to:
  incanter.stats/sample-normal, sample-uniform and Math/random  -> the injected streams of the tests
  incanter.core/sel m y x                                        -> data[y][x]  (map.clj:24)
  incanter.core/pow                                              -> Math/pow
  incanter.stats/median                                          -> the median
  geometry-msgs / std-msgs record constructors (keyword args with the defaults of
  src/geometry_msgs.clj:24,48,73: point 0 0 0, quaternion 0 0 0 1)  -> plain maps with the same keys
  taoensso.timbre logging, core.async                           -> no-ops (never on the evaluated path)
Numbers follow Clojure on the JVM: longs and doubles, `/` of two longs is a ratio, `=` does not equate a
long with a double, Math/round is floor(x + 0.5) to a long, `int` truncates.  java.lang.Math's
transcendental functions are stood in for by the C library's (<= 1 ulp apart: the golden test allows 1e-12).

Used by tests/golden/make_golden_from_reference.py (which writes tests/golden/amclj_reference_golden.json)
and by tests/test_reference_text.py.
"""
from __future__ import annotations

import math
import re
from fractions import Fraction
from functools import reduce
from pathlib import Path


# ------------------------------------------------------------------------------------------
# data
# ------------------------------------------------------------------------------------------
class Symbol(str):
    __slots__ = ()

    def __repr__(self):
        return f"'{str(self)}"


class Keyword:
    __slots__ = ("name",)
    _interned: dict = {}

    def __new__(cls, name):
        k = cls._interned.get(name)
        if k is None:
            k = object.__new__(cls)
            k.name = name
            cls._interned[name] = k
        return k

    def __repr__(self):
        return ":" + self.name

    def __call__(self, m, default=None):  # (:key m)
        if m is None:
            return default
        return m.get(self, default)


K = Keyword


class ListForm(list):
    """( ... ) in source."""


class VecForm(list):
    """[ ... ] in source; also the runtime vector."""


class MapForm(list):
    """{ ... } in source: flat key/value list (evaluated into a dict)."""


class Recur(Exception):
    def __init__(self, args):
        self.args_ = args


class Atom:
    def __init__(self, v):
        self.v = v


# ------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""\s*(?:,|;[^\n]*\n?)*\s*(                 # skip blanks, commas, comments
      \#_ | \#\( | \#\{ | \^ | ~@ | [\[\]{}()'`~@]               # punctuation
    | "(?:\\.|[^\\"])*"                                           # string
    | [^\s\[\]{}()"`,;]+                                          # atom (x', rot1* are symbols)
    )?""", re.X)

_INT = re.compile(r"^[+-]?\d+$")
_FLOAT = re.compile(r"^[+-]?(\d+\.\d*|\.\d+|\d+)([eE][+-]?\d+)?$")


def _tokens(text):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            if text[pos:].strip(" \n\t,") == "":
                break
            raise SyntaxError(f"cannot tokenise at {text[pos:pos + 40]!r}")
        pos = m.end()
        if m.group(1):
            out.append(m.group(1))
    return out


def _atom(tok):
    if tok[0] == '"':
        return bytes(tok[1:-1], "utf-8").decode("unicode_escape")
    if _INT.match(tok):
        return int(tok)
    if _FLOAT.match(tok):
        return float(tok)
    if tok == "nil":
        return None
    if tok == "true":
        return True
    if tok == "false":
        return False
    if tok[0] == ":" and len(tok) > 1:
        return Keyword(tok[1:])
    return Symbol(tok)


class _Reader:
    def __init__(self, text):
        self.t, self.i = _tokens(text), 0

    def more(self):
        return self.i < len(self.t)

    def read(self):
        tok = self.t[self.i]
        self.i += 1
        if tok == "(":
            return ListForm(self._seq(")"))
        if tok == "[":
            return VecForm(self._seq("]"))
        if tok == "{":
            return MapForm(self._seq("}"))
        if tok == "#{":
            return ListForm([Symbol("hash-set")] + self._seq("}"))
        if tok == "#(":
            body = ListForm(self._seq(")"))
            return ListForm([Symbol("fn*"), body])
        if tok == "#_":
            self.read()
            return self.read() if self.more() else None
        if tok == "^":
            self.read()  # metadata: dropped
            return self.read()
        if tok == "'":
            return ListForm([Symbol("quote"), self.read()])
        if tok == "@":
            return ListForm([Symbol("deref"), self.read()])
        if tok in (")", "]", "}"):
            raise SyntaxError("unbalanced " + tok)
        return _atom(tok)

    def _seq(self, close):
        out = []
        while self.t[self.i] != close:
            if self.t[self.i] == "#_":
                self.i += 1
                self.read()
                continue
            out.append(self.read())
        self.i += 1
        return out


def read_all(text):
    r, forms = _Reader(text), []
    while r.more():
        if r.t[r.i] == "#_":
            r.i += 1
            r.read()
            continue
        forms.append(r.read())
    return forms


# ------------------------------------------------------------------------------------------
# numbers the way Clojure on the JVM has them
# ------------------------------------------------------------------------------------------
def _norm(x):
    if isinstance(x, Fraction) and x.denominator == 1:
        return int(x)
    return x


def _is_int(x):
    return isinstance(x, int) and not isinstance(x, bool)


def clj_div(*a):
    if len(a) == 1:
        a = (1,) + a
    acc = a[0]
    for b in a[1:]:
        if isinstance(acc, float) or isinstance(b, float):
            acc = float(acc) / float(b) if b != 0 else (math.nan if acc == 0 or acc != acc else math.copysign(math.inf, float(acc)) * (math.copysign(1.0, float(b))))
        else:
            acc = _norm(Fraction(acc) / Fraction(b))
    return acc


def clj_eq(*a):
    def eq(x, y):
        if isinstance(x, bool) or isinstance(y, bool):
            return x is y
        nx, ny = isinstance(x, (int, float, Fraction)), isinstance(y, (int, float, Fraction))
        if nx and ny:  # (= 1 1.0) is false: different categories
            return isinstance(x, float) == isinstance(y, float) and x == y
        return x == y
    return all(eq(a[i], a[i + 1]) for i in range(len(a) - 1))


def java_round(x):
    if _is_int(x):
        return x
    x = float(x)
    if x != x:
        return 0
    return int(math.floor(x + 0.5))


def clj_int(x):
    return int(x) if not isinstance(x, Fraction) else int(x.numerator / x.denominator) if False else math.trunc(x)


def median(xs):
    s = sorted(xs)
    n = len(s)
    return s[n // 2] if n % 2 else (s[n // 2 - 1] + s[n // 2]) / 2.0


MATH = {
    "abs": abs, "cos": lambda x: math.cos(x), "sin": lambda x: math.sin(x), "exp": lambda x: _exp(x),
    "sqrt": lambda x: math.sqrt(x), "hypot": lambda a, b: math.hypot(a, b),
    "atan2": lambda a, b: math.atan2(a, b), "pow": lambda a, b: math.pow(a, b), "round": java_round,
    "floor": lambda x: float(math.floor(x)), "ceil": lambda x: float(math.ceil(x)), "PI": math.pi,
}


def _exp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


# ------------------------------------------------------------------------------------------
# evaluator
# ------------------------------------------------------------------------------------------
class Env:
    __slots__ = ("vars", "parent")

    def __init__(self, parent=None):
        self.vars, self.parent = {}, parent

    def find(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e
            e = e.parent
        return None


class Namespace:
    def __init__(self, name):
        self.name, self.vars, self.aliases, self.refers, self.refer_all = name, {}, {}, {}, []


class Fn:
    def __init__(self, interp, ns, arities, env, name=None):
        self.interp, self.ns, self.arities, self.env, self.name = interp, ns, arities, env, name

    def __call__(self, *args):
        params, rest, body = self._pick(len(args))
        while True:
            env = Env(self.env)
            self.interp.bind_params(params, rest, args, env, self.ns)
            try:
                out = None
                for f in body:
                    out = self.interp.eval(f, env, self.ns)
                return out
            except Recur as r:
                args = r.args_

    def _pick(self, n):
        for params, rest, body in self.arities:
            if (rest is None and len(params) == n) or (rest is not None and n >= len(params)):
                return params, rest, body
        raise TypeError(f"wrong number of args ({n}) passed to {self.name}")


_AMP = Symbol("&")


class Interp:
    def __init__(self):
        self.namespaces: dict[str, Namespace] = {}
        self.core = self._core()

    # -- namespaces ---------------------------------------------------------------------------
    def ns(self, name):
        if name not in self.namespaces:
            self.namespaces[name] = Namespace(name)
        return self.namespaces[name]

    def stub_ns(self, name, mapping):
        self.ns(name).vars.update(mapping)

    def load(self, path):
        """Evaluate every top-level form of a source file, in order (as `require` would)."""
        forms = read_all(Path(path).read_text(encoding="utf-8"))
        cur = self.ns("user")
        for f in forms:
            if isinstance(f, ListForm) and f and f[0] == "ns":
                cur = self._ns_form(f)
                continue
            self.eval(f, Env(), cur)
        return cur

    def _ns_form(self, f):
        name = next(x for x in f[1:] if isinstance(x, Symbol))
        ns = self.ns(str(name))
        for clause in f[2:]:
            if isinstance(clause, ListForm) and clause and clause[0] is K("require"):
                for spec in clause[1:]:
                    if isinstance(spec, Symbol):
                        self.ns(str(spec))
                        continue
                    target = str(spec[0])
                    self.ns(target)
                    opts = dict(zip(spec[1::2], spec[2::2]))
                    if K("as") in opts:
                        ns.aliases[str(opts[K("as")])] = target
                    ref = opts.get(K("refer"))
                    if ref is K("all"):
                        ns.refer_all.append(target)
                    elif ref:
                        for s in ref:
                            ns.refers[str(s)] = target
        return ns

    # -- symbol resolution ----------------------------------------------------------------------
    def resolve(self, sym, env, ns):
        e = env.find(sym)
        if e is not None:
            return e.vars[sym]
        s = str(sym)
        if "/" in s and s != "/":
            pre, name = s.split("/", 1)
            if pre == "Math":
                return MATH[name]
            target = ns.aliases.get(pre, pre)
            if target in self.namespaces and name in self.namespaces[target].vars:
                return self.namespaces[target].vars[name]
            raise NameError(f"unresolved {s} in {ns.name}")
        if s in ns.vars:
            return ns.vars[s]
        if s in ns.refers and s in self.namespaces[ns.refers[s]].vars:
            return self.namespaces[ns.refers[s]].vars[s]
        for t in ns.refer_all:
            if s in self.namespaces[t].vars:
                return self.namespaces[t].vars[s]
        if s in self.core:
            return self.core[s]
        raise NameError(f"unresolved symbol {s} in {ns.name}")

    # -- destructuring --------------------------------------------------------------------------
    def destructure(self, pat, val, env, ns):
        if isinstance(pat, Symbol):
            env.vars[pat] = val
        elif isinstance(pat, VecForm):
            seq = list(val) if val is not None else []
            i = 0
            j = 0
            while j < len(pat):
                p = pat[j]
                if p == _AMP:
                    self.destructure(pat[j + 1], seq[i:] or None, env, ns)
                    j += 2
                    continue
                if p is K("as"):
                    env.vars[pat[j + 1]] = val
                    j += 2
                    continue
                self.destructure(p, seq[i] if i < len(seq) else None, env, ns)
                i += 1
                j += 1
        elif isinstance(pat, MapForm):
            m = val
            if isinstance(val, (list, tuple)):  # keyword arguments: & {:keys [...]}
                m = dict(zip(val[0::2], val[1::2]))
            m = m or {}
            spec = dict(zip(pat[0::2], pat[1::2]))
            defaults = spec.pop(K("or"), None)
            dflt = dict(zip(defaults[0::2], defaults[1::2])) if defaults is not None else {}
            for k, v in spec.items():
                if k is K("keys"):
                    for s in v:
                        key = K(str(s))
                        env.vars[Symbol(str(s))] = m[key] if key in m else (self.eval(dflt[s], env, ns) if s in dflt else None)
                elif k is K("as"):
                    env.vars[v] = m
                else:  # {local :key}
                    got = m.get(v) if not isinstance(v, Keyword) else v(m)
                    if got is None and k in dflt:
                        got = self.eval(dflt[k], env, ns)
                    self.destructure(k, got, env, ns)
        else:
            raise SyntaxError(f"cannot destructure {pat!r}")

    def bind_params(self, params, rest, args, env, ns):
        for p, a in zip(params, args):
            self.destructure(p, a, env, ns)
        if rest is not None:
            self.destructure(rest, list(args[len(params):]) or None, env, ns)

    # -- fn construction --------------------------------------------------------------------------
    def _arity(self, params, body):
        ps, rest = list(params), None
        if _AMP in ps:
            k = ps.index(_AMP)
            ps, rest = ps[:k], params[k + 1]
        return ps, rest, list(body)

    def make_fn(self, forms, env, ns, name=None):
        """forms: after the optional name: [params] body...  or  ([params] body...)+"""
        if forms and isinstance(forms[0], str) and not isinstance(forms[0], Symbol):
            forms = forms[1:]  # docstring
        if forms and isinstance(forms[0], MapForm):
            forms = forms[1:]  # attr map
        if isinstance(forms[0], VecForm):
            arities = [self._arity(forms[0], forms[1:])]
        else:
            arities = [self._arity(a[0], a[1:]) for a in forms]
        return Fn(self, ns, arities, env, name)

    # -- eval ------------------------------------------------------------------------------------------
    def eval(self, f, env, ns):
        if isinstance(f, Symbol):
            return self.resolve(f, env, ns)
        if isinstance(f, VecForm):
            return VecForm(self.eval(x, env, ns) for x in f)
        if isinstance(f, MapForm):
            return {self.eval(k, env, ns): self.eval(v, env, ns) for k, v in zip(f[0::2], f[1::2])}
        if not isinstance(f, ListForm):
            return f
        if not f:
            return f
        head = f[0]
        if isinstance(head, Symbol):
            sf = _SPECIAL.get(str(head))
            if sf is not None and env.find(head) is None:
                return sf(self, f, env, ns)
        fn = self.eval(head, env, ns)
        args = [self.eval(x, env, ns) for x in f[1:]]
        if isinstance(fn, dict):
            return fn.get(args[0], args[1] if len(args) > 1 else None)
        if isinstance(fn, VecForm):
            return fn[args[0]]
        return fn(*args)

    # -- clojure.core ------------------------------------------------------------------------------------
    def _core(self):
        def assoc(m, *kv):
            if isinstance(m, VecForm):
                out = VecForm(m)
                for k, v in zip(kv[0::2], kv[1::2]):
                    if k == len(out):
                        out.append(v)
                    else:
                        out[k] = v
                return out
            out = dict(m or {})
            for k, v in zip(kv[0::2], kv[1::2]):
                out[k] = v
            return out

        def get(m, k, d=None):
            if m is None:
                return d
            if isinstance(m, dict):
                return m.get(k, d)
            if isinstance(m, (list, tuple)):
                return m[k] if _is_int(k) and 0 <= k < len(m) else d
            return d

        def get_in(m, ks, d=None):
            for k in ks:
                m = get(m, k, None)
                if m is None:
                    return d
            return m

        def update_in(m, ks, fn, *args):
            k = ks[0]
            if len(ks) == 1:
                return assoc(m, k, fn(get(m, k), *args))
            return assoc(m, k, update_in(get(m, k), ks[1:], fn, *args))

        def add(*a):
            return _norm(sum(a[1:], a[0])) if a else 0

        def sub(*a):
            if len(a) == 1:
                return -a[0]
            return _norm(reduce(lambda x, y: x - y, a))

        def mul(*a):
            return _norm(reduce(lambda x, y: x * y, a, 1))

        def cmp(op):
            return lambda *a: all(op(a[i], a[i + 1]) for i in range(len(a) - 1))

        def clj_range(*a):
            return VecForm(range(*a))

        def conj(c, *xs):
            if c is None:
                return VecForm(reversed(xs))
            return VecForm(list(c) + list(xs))

        def nth(c, i, *d):
            try:
                return list(c)[i] if i >= 0 else (_ for _ in ()).throw(IndexError())
            except IndexError:
                if d:
                    return d[0]
                raise

        def comp(*fs):
            def composed(*a):
                r = fs[-1](*a)
                for g in reversed(fs[:-1]):
                    r = g(r)
                return r
            return composed

        def swap(atom, fn, *args):
            atom.v = fn(atom.v, *args)
            return atom.v

        def reset(atom, v):
            atom.v = v
            return v

        def clj_reduce(fn, *a):
            if len(a) == 1:
                xs = list(a[0])
                return reduce(fn, xs) if xs else fn()
            return reduce(fn, list(a[1]), a[0])

        def clj_map(fn, *cs):
            return VecForm(fn(*xs) for xs in zip(*[list(c) for c in cs]))

        def clj_max(*a):
            return reduce(lambda x, y: y if y > x else x, a)

        def clj_min(*a):
            return reduce(lambda x, y: y if y < x else x, a)

        return {
            "+": add, "-": sub, "*": mul, "/": clj_div, "=": clj_eq, "==": cmp(lambda x, y: x == y),
            "not=": lambda *a: not clj_eq(*a),
            "<": cmp(lambda x, y: x < y), ">": cmp(lambda x, y: x > y), "<=": cmp(lambda x, y: x <= y),
            ">=": cmp(lambda x, y: x >= y), "inc": lambda x: x + 1, "dec": lambda x: x - 1,
            "max": clj_max, "min": clj_min, "zero?": lambda x: x == 0, "pos?": lambda x: x > 0,
            "neg?": lambda x: x < 0, "even?": lambda x: x % 2 == 0, "odd?": lambda x: x % 2 == 1,
            "nil?": lambda x: x is None, "not": lambda x: x is None or x is False,
            "int": clj_int, "long": clj_int, "double": float, "float": float, "abs": abs,
            "count": lambda c: 0 if c is None else len(c), "first": lambda c: (list(c) or [None])[0] if c is not None else None,
            "second": lambda c: (list(c) + [None, None])[1], "rest": lambda c: VecForm(list(c)[1:]),
            "next": lambda c: VecForm(list(c)[1:]) or None, "last": lambda c: (list(c) or [None])[-1],
            "nth": nth, "conj": conj, "vec": lambda c: VecForm(c or []), "vector": lambda *a: VecForm(a),
            "list": lambda *a: VecForm(a), "vector?": lambda x: isinstance(x, VecForm), "map?": lambda x: isinstance(x, dict),
            "seq": lambda c: (VecForm(c) if c else None), "empty?": lambda c: not c,
            "range": clj_range, "map": clj_map, "mapv": clj_map,
            "filter": lambda fn, c: VecForm(x for x in c if fn(x) not in (None, False)),
            "reduce": clj_reduce, "apply": lambda fn, *a: fn(*(list(a[:-1]) + list(a[-1] or []))),
            "take": lambda n, c: VecForm(list(c)[:n]), "drop": lambda n, c: VecForm(list(c)[n:]),
            "repeatedly": lambda n, fn: VecForm(fn() for _ in range(n)),
            "repeat": lambda n, x: VecForm([x] * n), "concat": lambda *cs: VecForm(x for c in cs for x in (c or [])),
            "reverse": lambda c: VecForm(reversed(list(c))), "sort": lambda c: VecForm(sorted(c)),
            "assoc": assoc, "dissoc": lambda m, *ks: {k: v for k, v in m.items() if k not in ks},
            "get": get, "get-in": get_in, "update-in": update_in, "assoc-in": lambda m, ks, v: update_in(m, ks, lambda _: v),
            "merge": lambda *ms: {k: v for m in ms if m for k, v in m.items()},
            "keys": lambda m: VecForm(m.keys()), "vals": lambda m: VecForm(m.values()),
            "comp": comp, "partial": lambda fn, *a: (lambda *b: fn(*a, *b)), "identity": lambda x: x,
            "constantly": lambda x: (lambda *a: x),
            "str": lambda *a: "".join("" if x is None else str(x) for x in a), "println": lambda *a: None,
            "atom": Atom, "deref": lambda a: a.v, "swap!": swap, "reset!": reset,
            "hash-set": lambda *a: set(a), "contains?": lambda m, k: k in m,
        }


# ------------------------------------------------------------------------------------------
# special forms
# ------------------------------------------------------------------------------------------
def _truthy(v):
    return not (v is None or v is False)


def _sf_def(i, f, env, ns):
    name = f[1]
    val = i.eval(f[-1], env, ns) if len(f) > 2 else None
    ns.vars[str(name)] = val
    return val


def _sf_defn(i, f, env, ns):
    name = f[1]
    fn = i.make_fn(f[2:], Env(), ns, str(name))
    ns.vars[str(name)] = fn
    return fn


def _sf_fn(i, f, env, ns):
    forms = f[1:]
    name = None
    if forms and isinstance(forms[0], Symbol):
        name, forms = forms[0], forms[1:]
    fn_env = Env(env)
    fn = i.make_fn(forms, fn_env, ns, str(name) if name else "fn")
    if name:
        fn_env.vars[name] = fn
    return fn


def _sf_fnstar(i, f, env, ns):  # #( ... ) with %, %1, %2, %&
    body = f[1]

    def scan(x, acc):
        if isinstance(x, Symbol) and x.startswith("%"):
            acc.add(str(x))
        elif isinstance(x, list):
            for y in x:
                scan(y, acc)
        return acc
    used = scan(body, set())
    n = max([1 if s == "%" else int(s[1:]) for s in used if s != "%&"] or [0])
    params = VecForm(Symbol(f"%{k}") for k in range(1, n + 1))
    if "%&" in used:
        params += [_AMP, Symbol("%&")]
    fn = Fn(i, ns, [i._arity(params, [body])], Env(env), "#()")
    orig = fn.__call__

    class P(Fn):
        def __call__(self, *args):
            params_, rest, body_ = self._pick(len(args))
            e = Env(self.env)
            i.bind_params(params_, rest, args, e, ns)
            if args:
                e.vars[Symbol("%")] = args[0]
            return i.eval(body_[0], e, ns)
    fn.__class__ = P
    return fn


def _sf_let(i, f, env, ns):
    e = Env(env)
    b = f[1]
    for k in range(0, len(b), 2):
        i.destructure(b[k], i.eval(b[k + 1], e, ns), e, ns)
    out = None
    for x in f[2:]:
        out = i.eval(x, e, ns)
    return out


def _sf_loop(i, f, env, ns):
    b = f[1]
    pats = b[0::2]
    vals = [None] * len(pats)
    e = Env(env)
    for k, p in enumerate(pats):
        vals[k] = i.eval(b[2 * k + 1], e, ns)
        i.destructure(p, vals[k], e, ns)
    while True:
        try:
            out = None
            for x in f[2:]:
                out = i.eval(x, e, ns)
            return out
        except Recur as r:
            e = Env(env)
            for p, v in zip(pats, r.args_):
                i.destructure(p, v, e, ns)


def _sf_recur(i, f, env, ns):
    raise Recur([i.eval(x, env, ns) for x in f[1:]])


def _sf_if(i, f, env, ns):
    if _truthy(i.eval(f[1], env, ns)):
        return i.eval(f[2], env, ns)
    return i.eval(f[3], env, ns) if len(f) > 3 else None


def _sf_cond(i, f, env, ns):
    for k in range(1, len(f) - 1, 2):
        t = f[k]
        if t is K("else") or _truthy(i.eval(t, env, ns)):
            return i.eval(f[k + 1], env, ns)
    return None


def _sf_do(i, f, env, ns):
    out = None
    for x in f[1:]:
        out = i.eval(x, env, ns)
    return out


def _sf_when(i, f, env, ns):
    return _sf_do(i, ListForm([None] + list(f[2:])), env, ns) if _truthy(i.eval(f[1], env, ns)) else None


def _sf_when_not(i, f, env, ns):
    return _sf_do(i, ListForm([None] + list(f[2:])), env, ns) if not _truthy(i.eval(f[1], env, ns)) else None


def _sf_and(i, f, env, ns):
    out = True
    for x in f[1:]:
        out = i.eval(x, env, ns)
        if not _truthy(out):
            return out
    return out


def _sf_or(i, f, env, ns):
    out = None
    for x in f[1:]:
        out = i.eval(x, env, ns)
        if _truthy(out):
            return out
    return out


def _thread(first):
    def sf(i, f, env, ns):
        x = f[1]
        for step in f[2:]:
            if isinstance(step, ListForm):
                x = ListForm([step[0], x] + list(step[1:])) if first else ListForm(list(step) + [x])
            else:
                x = ListForm([step, x])
        return i.eval(x, env, ns)
    return sf


def _sf_for(i, f, env, ns):
    out = VecForm()
    b = f[1]

    def rec(k, e):
        if k >= len(b):
            out.append(i.eval(f[2], e, ns))
            return
        pat, src = b[k], b[k + 1]
        if pat is K("when"):
            if _truthy(i.eval(src, e, ns)):
                rec(k + 2, e)
            return
        if pat is K("let"):
            e2 = Env(e)
            for q in range(0, len(src), 2):
                i.destructure(src[q], i.eval(src[q + 1], e2, ns), e2, ns)
            rec(k + 2, e2)
            return
        for v in list(i.eval(src, e, ns) or []):
            e2 = Env(e)
            i.destructure(pat, v, e2, ns)
            rec(k + 2, e2)
    rec(0, env)
    return out


def _sf_quote(i, f, env, ns):
    return f[1]


def _sf_assert(i, f, env, ns):
    if not _truthy(i.eval(f[1], env, ns)):
        raise AssertionError(str(f[2]) if len(f) > 2 else "assert failed")
    return None


def _sf_unsupported(i, f, env, ns):
    raise NotImplementedError(f"({f[0]} ...) is outside the evaluated subset (it is not on the hot path)")


def _sf_declare(i, f, env, ns):
    return None


_SPECIAL = {
    "def": _sf_def, "defn": _sf_defn, "defn-": _sf_defn, "fn": _sf_fn, "fn*": _sf_fnstar, "let": _sf_let,
    "loop": _sf_loop, "recur": _sf_recur, "if": _sf_if, "cond": _sf_cond, "do": _sf_do, "when": _sf_when,
    "when-not": _sf_when_not, "and": _sf_and, "or": _sf_or, "->": _thread(True), "->>": _thread(False),
    "for": _sf_for, "quote": _sf_quote, "assert": _sf_assert, "declare": _sf_declare, "comment": _sf_declare,
    "try": _sf_unsupported, "go": _sf_unsupported, "go-loop": _sf_unsupported, "doto": _sf_unsupported,
    "defrecord": _sf_declare, "defmethod": _sf_declare, "defmulti": _sf_declare, "defprotocol": _sf_declare,
}


# ------------------------------------------------------------------------------------------
# the reference, loaded
# ------------------------------------------------------------------------------------------
class Streams:
    """The injected noise / uniform streams (north_star: 'the same pre-generated noise/uniform streams
    injected').  normals: standard normals in draw order; uniform_ints: the integers sample-uniform
    :integers draws; randoms: Math/random doubles in [0, 1); uniforms: real sample-uniform draws as
    fractions in [0, 1)."""

    def __init__(self, normals=(), uniform_ints=(), randoms=(), uniforms=()):
        self.normals, self.uniform_ints, self.randoms, self.uniforms = (list(normals), list(uniform_ints), list(randoms),
                                                                        list(uniforms))
        self.used = {"normals": 0, "uniform_ints": 0, "randoms": 0, "uniforms": 0}

    def _pop(self, name):
        seq = getattr(self, name)
        k = self.used[name]
        if k >= len(seq):
            raise IndexError(f"injected stream '{name}' exhausted after {k} draws")
        self.used[name] = k + 1
        return seq[k]


def _kwargs(a):
    return dict(zip(a[0::2], a[1::2]))


def _record(defaults):
    def ctor(*a):
        kw = _kwargs(a)
        return {K(k): kw.get(K(k), d() if callable(d) else d) for k, d in defaults.items()}
    return ctor


def load_reference(src="/root/reference/src", streams: Streams | None = None):
    """The five hot-path namespaces, evaluated from the reference's own files.  Returns the interpreter."""
    it = Interp()
    st = streams or Streams()
    it.streams = st
    src = Path(src)

    def sample_normal(n, *a):
        kw = _kwargs(a)
        mean, sd = kw.get(K("mean"), 0), kw.get(K("sd"), 1)
        draw = lambda: mean + sd * it.streams._pop("normals")  # noqa: E731
        return draw() if n == 1 else VecForm(draw() for _ in range(n))  # scalar for n = 1: motion.clj:22 adds it

    def sample_uniform(n, *a):
        kw = _kwargs(a)
        lo, hi = kw.get(K("min"), 0.0), kw.get(K("max"), 1.0)
        if kw.get(K("integers")):
            draw = lambda: it.streams._pop("uniform_ints")  # noqa: E731  (already in {min..max})
        else:
            draw = lambda: lo + (hi - lo) * it.streams._pop("uniforms")  # noqa: E731
        return VecForm(draw() for _ in range(n))  # a seq: pf.clj:161 takes `first`, map.clj:63 destructures

    MATH["random"] = lambda: it.streams._pop("randoms")
    point = _record({"x": 0, "y": 0, "z": 0})                      # src/geometry_msgs.clj:24
    quaternion = _record({"x": 0, "y": 0, "z": 0, "w": 1})          # :48
    pose = _record({"position": point, "orientation": quaternion})  # :73
    gm = {"point": point, "quaternion": quaternion, "pose": pose,
          "vector3": _record({"x": 0, "y": 0, "z": 0}),
          "transform": _record({"translation": lambda: None, "rotation": quaternion}),
          "pose-stamped": _record({"header": None, "pose": pose}),
          "pose-array": _record({"header": None, "poses": lambda: VecForm()})}
    it.stub_ns("geometry-msgs", gm)
    it.stub_ns("std-msgs", {"header": _record({"seq": 0, "stamp": None, "frameId": ""})})
    for name in ("nav-msgs", "tf2-msgs", "clojure.core.async"):
        it.ns(name)
    it.stub_ns("incanter.core", {"sel": lambda m, y, x: m[y][x], "pow": lambda a, b: math.pow(a, b),
                                 "matrix": lambda *a: a[0]})
    it.stub_ns("incanter.stats", {"sample-normal": sample_normal, "sample-uniform": sample_uniform, "median": median})
    nop = lambda *a: None  # noqa: E731
    it.stub_ns("taoensso.timbre", {k: nop for k in ("trace", "debug", "info", "warn", "error")})
    for f in ("quaternions", "map", "motion", "sensor", "pf"):
        it.load(src / "amclj" / f"{f}.clj")
    return it


def call(it, qualified, *args):
    ns, name = qualified.split("/")
    return it.namespaces[ns].vars[name](*args)
