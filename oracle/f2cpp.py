#!/usr/bin/env python
"""f2cpp -- mechanical transliteration of the reference's dycore routines from Fortran to C++.

TEST INFRASTRUCTURE ONLY (see oracle/README_ref.md).  The image has no Fortran compiler, so the reference's own
implementation of the hot path cannot be compiled directly.  What CAN be done is to feed the reference's source text,
statement by statement, through a translator that knows nothing about meteorology: every assignment, loop bound, branch
and argument list of the emitted C++ comes from the tokens of
    /root/reference/src/core_atmosphere/dynamics/mpas_atm_time_integration.F  (the *_work routines, their wrappers,
        the pool-based routines) and
    /root/reference/src/framework/mpas_constants.F  (the physical constants)
as they lie there, after the same `cpp -traditional-cpp` pass the reference's Makefile applies.  The output goes to
oracle/_ref/ (git-ignored, never committed: it is derived from reference text) and is compiled into
oracle/_ref/libtiref.so, which tests/test_reference_pin.py compares the hand-written oracle against, routine by routine.

The Fortran subset handled is what those routines use: explicit-shape / assumed-shape / pointer arrays with arbitrary
lower bounds, automatic local arrays, statement functions, do / do while / if / select case, pool look-ups
(mpas_pool_get_array / _dimension / _config / _field), optional and keyword arguments, `**` with integer and real
exponents, and the numeric intrinsics.  Semantics that matter for bit equality and how they are kept:
  * column-major, 1-based (or declared lower bound) indexing: FArr<T>::operator() in f2cpp_rt.h;
  * evaluation order: every binary operation is emitted fully parenthesised in the parse order Fortran's precedence
    and left-to-right (right-to-left for **) associativity rules give; the C++ is compiled -ffp-contract=off;
  * default-kind real literals are RKIND in the reference's builds (-fdefault-real-8 in the double build): emitted
    through RL(), which appends `f` in the single build;
  * x**n with integer n: the repeated-squaring order of libgcc's __powidf2, which gfortran calls; x**y: std::pow;
  * sign(a, b), max/min with any number of arguments, merge, mod, abs, sqrt as Fortran defines them.
Anything outside the subset raises: nothing is silently skipped except OpenMP/OpenACC directives (comments) and calls to
timers.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

# --------------------------------------------------------------------------------------------- source -> statements


def load_statements(path, defines=()):
    """cpp -traditional-cpp (reference Makefile: CPP = cpp -P -traditional), comments stripped, continuations joined."""
    cmd = ["cpp", "-P", "-traditional-cpp"] + [f"-D{d}" for d in defines] + [path]
    text = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    out, cur = [], ""
    for raw in text.split("\n"):
        # the OpenMP directives the work routines rely on when several threads run them at once (mpas_atm_threading.F ranges):
        # kept as statements; every other directive (PARALLEL DO of the callers, OpenACC) is a comment
        m = re.match(r"^\s*!\$OMP\s+(END\s+MASTER|MASTER|BARRIER)\s*$", raw, re.I)
        if m and not cur:
            out.append("__omp_" + re.sub(r"\s+", "_", m.group(1).lower()))
            continue
        s, q = "", None
        for ch in raw:
            if q:
                s += ch
                if ch == q:
                    q = None
            elif ch in "\"'":
                q = ch
                s += ch
            elif ch == "!":
                break
            else:
                s += ch
        s = s.strip()
        if not s:
            continue
        if cur:
            if s.startswith("&"):
                s = s[1:].lstrip()
            cur += " " + s
        else:
            cur = s
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        # several statements on a line
        out.extend(split_semicolons(cur))
        cur = ""
    return out


def split_semicolons(s):
    parts, cur, q = [], "", None
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "\"'":
            q = ch
            cur += ch
        elif ch == ";":
            if cur.strip():
                parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


# --------------------------------------------------------------------------------------------- tokens and expressions
DOTTED = {".and.": "&&", ".or.": "||", ".not.": "!", ".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">",
          ".ge.": ">=", ".true.": "TRUE", ".false.": "FALSE", ".eqv.": "EQV", ".neqv.": "NEQV"}
TOK = re.compile(r"""
    (?P<dot>\.(?:and|or|not|eq|ne|lt|le|gt|ge|true|false|eqv|neqv)\.)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][+-]?\d+)?(?:_\w+)?)
  | (?P<name>[A-Za-z]\w*)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<op>\*\*|//|==|/=|<=|>=|=>|[-+*/(),:=<>%])
  | (?P<ws>\s+)
""", re.X | re.I)


def tokenize(s):
    toks, pos = [], 0
    while pos < len(s):
        m = TOK.match(s, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {s[pos:pos + 30]!r} in {s!r}")
        kind = m.lastgroup
        text = m.group(0)
        if kind == "num":
            # "1.eq." : the number ends before the dot of a dotted operator
            m2 = re.match(r"(\d+)(\.(?:and|or|not|eq|ne|lt|le|gt|ge|eqv|neqv)\.)", s[pos:], re.I)
            if m2:
                text = m2.group(1)
        pos += len(text)
        if kind == "ws":
            continue
        if kind == "dot":
            toks.append(("op", DOTTED[text.lower()]))
        elif kind == "name":
            toks.append(("name", text.lower()))
        else:
            toks.append((kind, text))
    return toks


class Parser:
    """Fortran expression grammar (F2003 R722 precedence): .eqv. < .or. < .and. < .not. < relational < // < +,- < *,/ < **."""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, text):
        if self.peek() == ("op", text):
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.accept(text):
            raise SyntaxError(f"expected {text!r}, got {self.peek()} in {self.t}")

    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        a = self.p_or()
        while self.peek() in (("op", "EQV"), ("op", "NEQV")):
            op = self.next()[1]
            a = ("bin", op, a, self.p_or())
        return a

    def p_or(self):
        a = self.p_and()
        while self.accept("||"):
            a = ("bin", "||", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.accept("&&"):
            a = ("bin", "&&", a, self.p_not())
        return a

    def p_not(self):
        if self.accept("!"):
            return ("un", "!", self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_cat()
        if self.peek()[0] == "op" and self.peek()[1] in ("==", "/=", "<", "<=", ">", ">="):
            op = self.next()[1]
            a = ("bin", op, a, self.p_cat())
        return a

    def p_cat(self):
        a = self.p_add()
        while self.accept("//"):
            a = ("bin", "//", a, self.p_add())
        return a

    def p_add(self):
        if self.peek() in (("op", "-"), ("op", "+")):
            op = self.next()[1]
            a = ("un", op, self.p_mul())
        else:
            a = self.p_mul()
        while self.peek() in (("op", "-"), ("op", "+")):
            op = self.next()[1]
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.next()[1]
            a = ("bin", op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.accept("**"):
            # right associative; the exponent may carry a unary sign: a ** -b
            if self.peek() in (("op", "-"), ("op", "+")):
                op = self.next()[1]
                b = ("un", op, self.p_pow())
            else:
                b = self.p_pow()
            a = ("bin", "**", a, b)
        return a

    def p_primary(self):
        kind, text = self.next()
        if kind == "num":
            return ("num", text)
        if kind == "str":
            q = text[0]
            return ("str", text[1:-1].replace(q + q, q))
        if kind == "op" and text in ("TRUE", "FALSE"):
            return ("log", text == "TRUE")
        if kind == "op" and text == "(":
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        if kind == "name":
            node = ("name", text)
            while True:
                if self.accept("("):
                    args = self.arglist()
                    node = ("call", node, args)
                elif self.accept("%"):
                    k2, t2 = self.next()
                    assert k2 == "name"
                    node = ("member", node, t2)
                else:
                    break
            return node
        raise SyntaxError(f"unexpected token {kind} {text!r} in {self.t}")

    def arglist(self):
        args = []
        if self.accept(")"):
            return args
        while True:
            args.append(self.arg())
            if self.accept(")"):
                return args
            self.expect(",")

    def arg(self):
        # keyword argument, section or expression
        if self.peek()[0] == "name" and self.peek(1) == ("op", "=") :
            name = self.next()[1]
            self.next()
            return ("kw", name, self.expr())
        lo = None
        if self.peek() != ("op", ":"):
            lo = self.expr()
        if self.accept(":"):
            hi = None
            if self.peek() not in (("op", ","), ("op", ")")):
                hi = self.expr()
            return ("section", lo, hi)
        return lo


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.peek()[0] != "eof":
        raise SyntaxError(f"trailing tokens in expression {s!r}: {p.t[p.i:]}")
    return e


def split_top(s, sep=","):
    parts, cur, depth, q = [], "", 0, None
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
            continue
        if ch in "\"'":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip() or parts:
        parts.append(cur.strip())
    return parts


# --------------------------------------------------------------------------------------------- symbols
CPP_RESERVED = {"int", "float", "double", "new", "delete", "class", "register", "this", "template", "default", "auto", "and", "or",
                "not", "union", "struct", "char", "long", "short", "signed", "unsigned", "switch", "case", "break", "continue",
                "return", "if", "else", "for", "while", "do", "const", "static", "operator", "namespace", "using", "bool", "true",
                "false", "void", "real", "min", "max", "abs"}


def cname(n):
    n = n.lower()
    return n + "_" if n in CPP_RESERVED else n


class Sym:
    def __init__(self, name, base, dims=None, intent=None, pointer=False, optional=False, parameter=False, init=None,
                 dummy=False, typename=None, module=False):
        self.name, self.base, self.dims, self.intent = name, base, dims, intent
        self.pointer, self.optional, self.parameter, self.init = pointer, optional, parameter, init
        self.dummy, self.typename, self.module = dummy, typename, module
        self.stmt_fn = None            # (arg names, body expr) for statement functions

    @property
    def rank(self):
        return len(self.dims) if self.dims else 0

    @property
    def ctype(self):
        return {"real": "real", "integer": "int", "logical": "bool", "character": "std::string", "double": "double"}.get(self.base)


DECL = re.compile(r"^(real|integer|logical|character|type|procedure|double\s+precision)\b", re.I)


def parse_decl(stmt, dummies, module=False):
    """One type-declaration statement -> [Sym]."""
    if "::" in stmt:
        left, right = stmt.split("::", 1)
    else:
        m = re.match(r"^(\w+(?:\s*\([^)]*\))?)\s+(.*)$", stmt)
        left, right = m.group(1), m.group(2)
    parts = split_top(left)
    tspec = parts[0].strip()
    tl = tspec.lower().replace(" ", "")
    typename = None
    if tl.startswith("real"):
        base = "double" if "r8kind" in tl or "kind=8" in tl else "real"
    elif tl.startswith("doubleprecision"):
        base = "double"
    elif tl.startswith("integer"):
        base = "integer"
    elif tl.startswith("logical"):
        base = "logical"
    elif tl.startswith("character"):
        base = "character"
    elif tl.startswith("type"):
        base = "type"
        typename = re.match(r"type\((\w+)\)", tl).group(1)
    elif tl.startswith("procedure"):
        base = "procedure"
    else:
        raise SyntaxError(f"type spec {tspec!r}")
    attrs = dict(dims=None, intent=None, pointer=False, optional=False, parameter=False)
    for a in parts[1:]:
        al = a.lower().replace(" ", "")
        if al.startswith("dimension"):
            attrs["dims"] = parse_dims(a[a.index("(") + 1:a.rindex(")")])
        elif al.startswith("intent"):
            attrs["intent"] = al[al.index("(") + 1:-1]
        elif al in ("pointer", "allocatable"):
            attrs["pointer"] = True
        elif al == "optional":
            attrs["optional"] = True
        elif al == "parameter":
            attrs["parameter"] = True
        elif al in ("target", "private", "public", "protected", "save", "contiguous"):
            pass
        else:
            raise SyntaxError(f"attribute {a!r} in {stmt!r}")
    syms = []
    for ent in split_top(right):
        init = None
        m = re.match(r"^(\w+)\s*(\((.*)\))?\s*(?:=(?!>)\s*(.*))?$", ent.strip(), re.S)
        if not m:
            raise SyntaxError(f"entity {ent!r} in {stmt!r}")
        name = m.group(1).lower()
        dims = parse_dims(m.group(3)) if m.group(2) else attrs["dims"]
        if m.group(4) is not None:
            init = parse_expr(m.group(4))
        syms.append(Sym(name, base, dims, attrs["intent"], attrs["pointer"], attrs["optional"], attrs["parameter"], init,
                        name in dummies, typename, module))
    return syms


def parse_dims(s):
    dims = []
    for d in split_top(s):
        d = d.strip()
        if d == ":":
            dims.append((None, None))
        elif ":" in split_colon(d)[0] or len(split_colon(d)) == 2:
            lo, hi = split_colon(d)
            dims.append((parse_expr(lo) if lo.strip() else None, parse_expr(hi) if hi.strip() else None))
        else:
            dims.append((("num", "1"), parse_expr(d)))
    return dims


def split_colon(d):
    parts = split_top(d, ":")
    return parts if len(parts) == 2 else [d]


# --------------------------------------------------------------------------------------------- translation unit
INTRINSIC_REAL = {"sqrt", "exp", "log", "abs", "sign", "max", "min", "mod", "real", "merge", "dble", "float", "tanh", "cos", "sin",
                  "atan", "acos", "asin", "int", "nint", "floor", "trim", "present", "epsilon", "huge", "tiny", "size", "associated",
                  "amax1", "amin1", "len_trim"}
IGNORED_CALLS = {"mpas_timer_start", "mpas_timer_stop", "mpas_allocate_scratch_field", "mpas_deallocate_scratch_field"}


class Routine:
    def __init__(self, name, args, stmts):
        self.name, self.args, self.stmts = name, args, stmts
        self.syms = {}


class Unit:
    def __init__(self):
        self.module_syms = {}          # name -> Sym (module variables and parameters)
        self.routines = {}             # name -> Routine (parsed)
        self.out = []

    # ---- parsing of files
    def add_file(self, path, defines, want, module_spec=True):
        stmts = load_statements(path, defines)
        i, in_contains, depth_iface = 0, False, 0
        while i < len(stmts):
            s = stmts[i]
            sl = s.lower()
            if re.match(r"^(abstract\s+)?interface\b", sl):
                depth_iface += 1
            elif re.match(r"^end\s+interface", sl):
                depth_iface -= 1
            elif depth_iface:
                pass
            elif sl == "contains":
                in_contains = True
            elif re.match(r"^(recursive\s+)?subroutine\s+\w+", sl):
                m = re.match(r"^(?:recursive\s+)?subroutine\s+(\w+)\s*(?:\((.*)\))?\s*$", s, re.I | re.S)
                name = m.group(1).lower()
                args = [a.strip().lower() for a in split_top(m.group(2) or "") if a.strip()]
                j = i + 1
                body = []
                while not re.match(r"^end\s*subroutine", stmts[j].lower()):
                    body.append(stmts[j])
                    j += 1
                if name in want:
                    self.routines[name] = Routine(name, args, body)
                i = j
            elif not in_contains and module_spec and DECL.match(s) and "function" not in sl.split("::")[0]:
                try:
                    for sym in parse_decl(s, (), module=True):
                        self.module_syms.setdefault(sym.name, sym)
                except SyntaxError:
                    pass                   # declarations of derived types / clocks this subset never touches
            i += 1

    # ---- emission
    def emit_all(self, order):
        for sym in self.module_syms.values():
            self.emit_module_sym(sym)
        self.out.append("")
        for name in order:
            self.routines[name].syms = self.collect_syms(self.routines[name])
        for name in order:
            self.out.append(self.prototype(self.routines[name]) + ";")
        self.out.append("")
        for name in order:
            self.emit_routine(self.routines[name])
        return "\n".join(self.out) + "\n"

    def emit_module_sym(self, s):
        if s.base in ("type", "procedure"):
            return
        if s.parameter:
            if s.base == "integer" and s.name in ("rkind", "r8kind", "r4kind", "strkind", "shortstrkind"):
                return
            em = Emitter(self, None)
            self.out.append(f"static const {s.ctype} {cname(s.name)} = {em.conv(s.init, s.base)};")
        elif s.rank:
            self.out.append(f"FArr<{s.ctype}> {cname(s.name)};          // module array, bound by the harness")
        else:
            self.out.append(f"{s.ctype} {cname(s.name)}{{}};           // module scalar")

    def collect_syms(self, r):
        syms = {}
        for s in r.stmts:
            if DECL.match(s) and not re.match(r"^\w+\s*(\(.*\))?\s*=[^=>]", s):
                for sym in parse_decl(s, set(r.args)):
                    syms[sym.name] = sym
            elif re.match(r"^(use|implicit)\b", s, re.I):
                continue
            else:
                # statement functions sit between the declarations and the first executable statement
                m = re.match(r"^(\w+)\s*\(([^()]*)\)\s*=(?!=)\s*(.*)$", s)
                if m and m.group(1).lower() in syms and not syms[m.group(1).lower()].rank and syms[m.group(1).lower()].base in ("real", "integer", "double") \
                        and all(re.match(r"^\w+$", a.strip()) for a in m.group(2).split(",")):
                    f = syms[m.group(1).lower()]
                    if f.stmt_fn is None and not f.dummy:
                        f.stmt_fn = ([a.strip().lower() for a in m.group(2).split(",")], parse_expr(m.group(3)))
                        f.stmt_fn_order = len([1 for t in syms.values() if t.stmt_fn])
                        continue
                break
        for a in r.args:
            if a not in syms:
                raise SyntaxError(f"{r.name}: dummy {a} not declared")
        return syms

    def param_decl(self, r, a):
        s = r.syms[a]
        n = cname(a)
        if s.base == "type":
            return {"mpas_pool_type": f"Pool& {n}", "block_type": f"BlockT& {n}", "domain_type": f"DomainT& {n}"}[s.typename]
        if s.base == "procedure":
            return f"ExchFn {n}"
        if s.base == "character":
            return f"const std::string& {n}"
        if s.rank:
            return f"FArr<{s.ctype}> {n}__a"
        if s.optional:
            return f"Opt<{s.ctype}> {n}"
        if s.intent == "in" or not self.is_assigned(r, a):
            return f"const {s.ctype} {n}"
        return f"{s.ctype}& {n}"

    def is_assigned(self, r, a):
        # a scalar dummy without intent is passed by reference only if the routine itself assigns it
        pat = re.compile(rf"^(if\s*\(.*\)\s*)?{a}\s*=(?!=)", re.I)
        return r.syms[a].intent in ("out", "inout") or any(pat.match(s) for s in r.stmts)

    def prototype(self, r):
        ps = []
        for a in r.args:
            d = self.param_decl(r, a)
            if r.syms[a].optional:
                d += " = {}"
            ps.append(d)
        return f"void {cname(r.name)}({', '.join(ps)})"

    def emit_routine(self, r):
        em = Emitter(self, r)
        # default arguments belong on the prototype only
        ps = [self.param_decl(r, a) for a in r.args]
        o = [f"void {cname(r.name)}({', '.join(ps)}) {{"]
        # dummies: re-bind explicit-shape arrays to the declared extents (sequence association)
        for a in r.args:
            s = r.syms[a]
            if s.rank and s.base != "type":
                n = cname(a)
                if all(lo is None and hi is None for lo, hi in s.dims):
                    o.append(f"    FArr<{s.ctype}> {n} = {n}__a.rebased();")
                else:
                    o.append(f"    FArr<{s.ctype}> {n}; {n}.bind({n}__a.p{em.bounds(s.dims)});")
        # locals
        for name, s in r.syms.items():
            if s.dummy or s.stmt_fn:
                continue
            n = cname(name)
            if s.base == "type":
                t = {"field3dreal": "FieldT<real>", "field2dreal": "FieldT<real>", "field1dreal": "FieldT<real>",
                     "mpas_pool_type": "Pool*", "block_type": "BlockT*"}.get(s.typename)
                if t is None:
                    raise SyntaxError(f"{r.name}: local of type {s.typename}")
                o.append(f"    {t} {n}{{}};")
            elif s.base == "procedure":
                continue
            elif s.parameter:
                o.append(f"    const {s.ctype} {n} = {em.conv(s.init, s.base)};")
            elif s.rank and not s.pointer:
                o.append(f"    std::vector<{s.ctype}> {n}__buf(FArr<{s.ctype}>::count({em.bounds(s.dims)[2:]})); FArr<{s.ctype}> {n}; {n}.bind({n}__buf.data(){em.bounds(s.dims)});")
            elif s.rank:
                o.append(f"    FArr<{s.ctype}> {n};")
            else:
                init = f" = {em.conv(s.init, s.base)}" if s.init is not None else "{}"
                o.append(f"    {s.ctype} {n}{init};")
        for name, s in sorted(((n, t) for n, t in r.syms.items() if t.stmt_fn), key=lambda kv: kv[1].stmt_fn_order):
            if s.stmt_fn:
                args, body = s.stmt_fn
                em.local_scalars = {a: r.syms[a].base for a in args}
                ps2 = ", ".join(f"{r.syms[a].ctype} {cname(a)}" for a in args)
                o.append(f"    auto {cname(name)} = [&]({ps2}) -> {s.ctype} {{ return {em.conv(body, s.base)}; }};")
                em.local_scalars = {}
        o.extend(em.body())
        o.append("}")
        o.append("")
        self.out.extend(o)


class Emitter:
    def __init__(self, unit, routine):
        self.u, self.r = unit, routine
        self.local_scalars = {}
        self.ind = 1
        self.sel_stack = []

    # ---- symbols and types
    def sym(self, name):
        if self.r and name in self.r.syms:
            return self.r.syms[name]
        return self.u.module_syms.get(name)

    def typeof(self, e):
        k = e[0]
        if k == "num":
            t = e[1].lower()
            if re.match(r"^\d+(_\w+)?$", t):
                return "integer"
            if "d" in t.split("_")[0] or t.endswith("_r8kind"):
                return "double"
            return "real"
        if k == "str":
            return "character"
        if k == "log":
            return "logical"
        if k == "paren":
            return self.typeof(e[1])
        if k == "name":
            if e[1] in self.local_scalars:
                return self.local_scalars[e[1]]
            s = self.sym(e[1])
            if s is None:
                raise SyntaxError(f"{self.r.name if self.r else '<module>'}: unknown name {e[1]}")
            return s.base
        if k == "member":
            return "real"
        if k == "un":
            return "logical" if e[1] == "!" else self.typeof(e[2])
        if k == "bin":
            op = e[1]
            if op in ("&&", "||", "EQV", "NEQV", "==", "/=", "<", "<=", ">", ">="):
                return "logical"
            if op == "//":
                return "character"
            a, b = self.typeof(e[2]), self.typeof(e[3])
            if op == "**":
                return a if b == "integer" else self.promote(a, b)
            return self.promote(a, b)
        if k == "call":
            if e[1][0] == "member":
                return "real"
            n = e[1][1]
            s = self.sym(n)
            if s is not None and (s.rank or s.stmt_fn):
                return s.base
            if n in ("real", "float", "sqrt", "exp", "log", "tanh", "cos", "sin", "atan", "acos", "asin", "epsilon", "huge", "tiny", "amax1", "amin1"):
                if n == "real" and len(e[2]) > 1:
                    kk = e[2][1]
                    kk = kk[2] if kk[0] == "kw" else kk
                    return "double" if kk == ("name", "r8kind") else "real"
                return "real" if n in ("real", "float", "epsilon", "huge", "tiny") else self.typeof(e[2][0])
            if n == "dble":
                return "double"
            if n in ("int", "nint", "floor", "size", "len_trim"):
                return "integer"
            if n in ("present", "associated"):
                return "logical"
            if n == "trim":
                return "character"
            if n in ("abs", "sign", "mod"):
                return self.typeof(e[2][0])
            if n in ("max", "min"):
                t = self.typeof(e[2][0])
                for a in e[2][1:]:
                    t = self.promote(t, self.typeof(a))
                return t
            if n == "merge":
                return self.promote(self.typeof(e[2][0]), self.typeof(e[2][1]))
            raise SyntaxError(f"{self.r.name}: type of call to {n}")
        raise SyntaxError(f"typeof {e}")

    @staticmethod
    def promote(a, b):
        order = ["logical", "integer", "real", "double"]
        if a == "character" or b == "character":
            return "character"
        return order[max(order.index(a), order.index(b))]

    # ---- expressions
    def conv(self, e, want=None):
        return self.ex(e)

    def lit(self, text):
        t = text.lower()
        m = re.match(r"^(.*?)(?:_(\w+))?$", t)
        body, kind = m.group(1), m.group(2)
        if re.match(r"^\d+$", body):
            return body
        if "d" in body:
            return "(double)" + body.replace("d", "e")
        if kind == "r8kind":
            return "(double)" + body
        if "." not in body and "e" not in body:
            body += ".0"
        if "." not in body:
            body = body.replace("e", ".0e")
        return f"RL({body})"

    def ex(self, e):
        k = e[0]
        if k == "num":
            return self.lit(e[1])
        if k == "str":
            return 'std::string("' + e[1].replace("\\", "\\\\").replace('"', '\\"') + '")'
        if k == "log":
            return "true" if e[1] else "false"
        if k == "paren":
            return "(" + self.ex(e[1]) + ")"
        if k == "name":
            s = self.sym(e[1])
            if s is not None and s.optional and not s.rank and s.base not in ("type", "procedure"):
                return cname(e[1]) + ".v"
            return cname(e[1])
        if k == "member":
            return self.ex(e[1]) + "." + cname(e[2])
        if k == "un":
            return f"({e[1]}{self.ex(e[2])})"
        if k == "bin":
            op, a, b = e[1], e[2], e[3]
            if op == "**":
                tb = self.typeof(b)
                if tb == "integer":
                    return f"f_powi({self.ex(a)}, {self.ex(b)})"
                return f"f_pow({self.ex(a)}, {self.ex(b)})"
            if op == "//":
                return f"({self.ex(a)} + {self.ex(b)})"
            if op == "/=":
                op = "!="
            if op == "EQV":
                op = "=="
            if op == "NEQV":
                op = "!="
            return f"({self.ex(a)} {op} {self.ex(b)})"
        if k == "call":
            if e[1][0] == "member":
                return self.ex(e[1]) + "(" + ", ".join(self.ex(a) for a in e[2]) + ")"
            n = e[1][1]
            s = self.sym(n)
            if s is not None and s.rank:
                if any(a[0] == "section" for a in e[2]):
                    # conformable section on the right-hand side of an array assignment: its k-th sectioned dimension runs with
                    # the k-th loop of the assignment (F2003 7.4.1.3), offset by the two lower bounds
                    loops = getattr(self, "section_loops", None)
                    if not loops:
                        raise SyntaxError(f"{self.r.name}: array section of {n} outside an array assignment")
                    idx, kth = [], 0
                    for d, a in enumerate(e[2]):
                        if a[0] == "section":
                            v, lo_l = loops[kth]
                            kth += 1
                            lo_r = self.ex(a[1]) if a[1] is not None else f"{cname(n)}.lo[{d}]"
                            idx.append(f"({lo_r} + ({v} - ({lo_l})))")
                        else:
                            idx.append(self.ex(a))
                    return f"{cname(n)}(" + ", ".join(idx) + ")"
                return f"{cname(n)}(" + ", ".join(self.ex(a) for a in e[2]) + ")"
            if s is not None and s.stmt_fn:
                return f"{cname(n)}(" + ", ".join(self.ex(a) for a in e[2]) + ")"
            args = [a[2] if a[0] == "kw" else a for a in e[2]]
            if n in ("real", "float"):
                t = self.typeof(e)
                return f"(({'double' if t == 'double' else 'real'})({self.ex(args[0])}))"
            if n == "dble":
                return f"((double)({self.ex(args[0])}))"
            if n == "int":
                return f"((int)({self.ex(args[0])}))"
            if n == "present":
                s0 = self.sym(args[0][1])
                return f"({cname(args[0][1])}.p != nullptr)" if s0.rank else f"{cname(args[0][1])}.present"
            if n == "associated":
                return f"f_associated({self.ex(args[0])})"
            if n == "trim":
                return self.ex(args[0])
            if n in INTRINSIC_REAL:
                return f"f_{n}(" + ", ".join(self.ex(a) for a in args) + ")"
            raise SyntaxError(f"{self.r.name}: unknown function {n}")
        raise SyntaxError(f"ex {e}")

    def bounds(self, dims):
        out = ""
        for lo, hi in dims:
            out += f", {self.ex(lo) if lo is not None else '1'}, {self.ex(hi)}"
        return out

    # ---- statements
    def line(self, s):
        return "    " * self.ind + s

    def body(self):
        o = []
        r = self.r
        started = False
        for s in r.stmts:
            if not started:
                if (DECL.match(s) and not re.match(r"^\w+\s*(\(.*\))?\s*=[^=>]", s)) or re.match(r"^(use|implicit)\b", s, re.I):
                    continue
                m = re.match(r"^(\w+)\s*\(", s)
                if m and m.group(1).lower() in r.syms and r.syms[m.group(1).lower()].stmt_fn and re.match(r"^\w+\s*\([^()]*\)\s*=(?!=)", s):
                    continue
                started = True
            o.extend(self.stmt(s))
        return o

    def stmt(self, s):
        sl = s.lower()
        o = []
        m = re.match(r"^if\s*\((.*)\)\s*then$", s, re.I | re.S)
        if m:
            o.append(self.line(f"if ({self.ex(parse_expr(m.group(1)))}) {{"))
            self.ind += 1
            return o
        m = re.match(r"^else\s*if\s*\((.*)\)\s*then$", s, re.I | re.S)
        if m:
            self.ind -= 1
            o.append(self.line(f"}} else if ({self.ex(parse_expr(m.group(1)))}) {{"))
            self.ind += 1
            return o
        if sl == "else":
            self.ind -= 1
            o.append(self.line("} else {"))
            self.ind += 1
            return o
        if re.match(r"^end\s*(if|do)$", sl):
            self.ind -= 1
            o.append(self.line("}"))
            return o
        m = re.match(r"^do\s+while\s*\((.*)\)$", s, re.I | re.S)
        if m:
            o.append(self.line(f"while ({self.ex(parse_expr(m.group(1)))}) {{"))
            self.ind += 1
            return o
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", s, re.I | re.S)
        if m:
            var = cname(m.group(1))
            parts = split_top(m.group(2))
            lo, hi = self.ex(parse_expr(parts[0])), self.ex(parse_expr(parts[1]))
            if len(parts) == 3:
                st = self.ex(parse_expr(parts[2]))
                o.append(self.line(f"for ({var} = {lo}; ({st}) > 0 ? {var} <= {hi} : {var} >= {hi}; {var} += {st}) {{"))
            else:
                o.append(self.line(f"for ({var} = {lo}; {var} <= {hi}; {var}++) {{"))
            self.ind += 1
            return o
        if sl == "do":
            o.append(self.line("while (true) {"))
            self.ind += 1
            return o
        m = re.match(r"^select\s*case\s*\((.*)\)$", s, re.I | re.S)
        if m:
            o.append(self.line(f"{{ const auto sel__ = {self.ex(parse_expr(m.group(1)))};"))
            self.ind += 1
            self.sel_stack.append(0)
            return o
        m = re.match(r"^case\s*\((.*)\)$", s, re.I | re.S)
        if m:
            conds = " || ".join(f"sel__ == {self.ex(parse_expr(v))}" for v in split_top(m.group(1)))
            if self.sel_stack[-1]:
                self.ind -= 1
                o.append(self.line(f"}} else if ({conds}) {{"))
            else:
                o.append(self.line(f"if ({conds}) {{"))
            self.sel_stack[-1] += 1
            self.ind += 1
            return o
        if re.match(r"^case\s+default$", sl):
            if self.sel_stack[-1]:
                self.ind -= 1
                o.append(self.line("} else {"))
            else:
                o.append(self.line("{"))
            self.sel_stack[-1] += 1
            self.ind += 1
            return o
        if re.match(r"^end\s*select$", sl):
            if self.sel_stack.pop():
                self.ind -= 1
                o.append(self.line("}"))
            self.ind -= 1
            o.append(self.line("}"))
            return o
        m = re.match(r"^if\s*\(", s, re.I)
        if m:
            # one-line if: find the matching parenthesis
            depth, j = 0, s.index("(")
            for j in range(s.index("("), len(s)):
                if s[j] == "(":
                    depth += 1
                elif s[j] == ")":
                    depth -= 1
                    if depth == 0:
                        break
            cond, rest = s[s.index("(") + 1:j], s[j + 1:].strip()
            o.append(self.line(f"if ({self.ex(parse_expr(cond))}) {{"))
            self.ind += 1
            o.extend(self.stmt(rest))
            self.ind -= 1
            o.append(self.line("}"))
            return o
        if sl == "__omp_barrier":
            return ["#pragma omp barrier"]
        if sl == "__omp_master":
            self.ind += 1
            return ["#pragma omp master", self.line("{")[4:]]
        if sl == "__omp_end_master":
            self.ind -= 1
            return [self.line("}")]
        if sl == "return":
            return [self.line("return;")]
        if sl == "cycle":
            return [self.line("continue;")]
        if sl == "exit":
            return [self.line("break;")]
        if sl == "continue":
            return []
        m = re.match(r"^call\s+(\w+)\s*(?:\((.*)\))?$", s, re.I | re.S)
        if m:
            return self.call(m.group(1).lower(), m.group(2) or "")
        # pointer assignment
        m = re.match(r"^([\w%\s]+?)\s*=>\s*(.*)$", s, re.S)
        if m:
            return [self.line(f"{self.ex(parse_expr(m.group(1)))} = {self.ex(parse_expr(m.group(2)))};")]
        # assignment: split at the top-level '='
        depth, q = 0, None
        for j, ch in enumerate(s):
            if q:
                if ch == q:
                    q = None
                continue
            if ch in "\"'":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and s[j + 1:j + 2] != "=" and s[j - 1] not in "<>/=":
                return self.assign(s[:j].strip(), s[j + 1:].strip())
        raise SyntaxError(f"{self.r.name}: statement not understood: {s!r}")

    def assign(self, lhs_s, rhs_s):
        lhs, rhs = parse_expr(lhs_s), parse_expr(rhs_s)
        if lhs[0] == "name":
            s = self.sym(lhs[1])
            if s is not None and s.rank:            # whole-array assignment of a scalar expression
                return [self.line(f"{cname(lhs[1])}.fill({self.ex(rhs)});")]
            return [self.line(f"{self.ex(lhs)} = {self.ex(rhs)};")]
        if lhs[0] == "call" and any(a[0] == "section" for a in lhs[2]):
            # a(:, i) = scalar expression  ->  loops over the sectioned dimensions
            n = lhs[1][1]
            s = self.sym(n)
            o, idx, loops = [], [], []
            # Fortran evaluates the whole right-hand side before storing; the loops below store as they go, which is the same
            # unless the right-hand side reads OTHER elements of the array being assigned -- none of the statements here does
            for d, a in reversed(list(enumerate(lhs[2]))):          # outermost loop = last dimension (column-major order)
                if a[0] == "section":
                    v = f"i{d}__"
                    lo = self.ex(a[1]) if a[1] is not None else f"{cname(n)}.lo[{d}]"
                    hi = self.ex(a[2]) if a[2] is not None else f"({cname(n)}.lo[{d}] + {cname(n)}.n[{d}] - 1)"
                    o.append(self.line(f"for (long {v} = {lo}; {v} <= {hi}; {v}++)"))
                    loops.insert(0, (v, lo))
            for d, a in enumerate(lhs[2]):
                idx.append(f"i{d}__" if a[0] == "section" else self.ex(a))
            self.section_loops = loops
            o.append(self.line(f"    {cname(n)}({', '.join(idx)}) = {self.ex(rhs)};"))
            self.section_loops = None
            return o
        return [self.line(f"{self.ex(lhs)} = {self.ex(rhs)};")]

    def call(self, name, argstr):
        p = Parser(tokenize("f(" + argstr + ")"))
        args = p.expr()[2]
        if name in IGNORED_CALLS:
            return []
        if name == "mpas_pool_get_array":
            var = args[2][1]
            s = self.sym(var)
            lev = self.ex(args[3][2] if args[3][0] == "kw" else args[3]) if len(args) > 3 else "1"
            if not s.rank:                 # a 0-d field (cf1, cf2, cf3): the pointer target is a scalar
                return [self.line(f"{cname(var)} = {self.ex(args[0])}.scalar<{s.ctype}>({self.ex(args[1])});")]
            return [self.line(f"{cname(var)} = {self.ex(args[0])}.arr<{s.ctype}>({self.ex(args[1])}, {lev}, {s.rank});")]
        if name == "mpas_pool_get_field":
            var = args[2][1]
            return [self.line(f"{cname(var)}.array = {self.ex(args[0])}.arr<real>({self.ex(args[1])}, 1, 0);")]
        if name == "mpas_pool_get_dimension":
            return [self.line(f"{cname(args[2][1])} = {self.ex(args[0])}.dim({self.ex(args[1])});")]
        if name == "mpas_pool_get_config":
            return [self.line(f"{self.ex(args[0])}.cfg({self.ex(args[1])}, {cname(args[2][1])});")]
        if name == "mpas_log_write":
            return [self.line(f"rt_log({self.ex(args[0])});")]
        s = self.sym(name)
        if s is not None and s.base == "procedure":            # exchange_halo_group(domain, group)
            return [self.line(f"{cname(name)}({self.ex(args[1])});")]
        if name not in self.u.routines:
            raise SyntaxError(f"{self.r.name}: call to untranslated routine {name}")
        callee = self.u.routines[name]
        actual = {}
        pos = 0
        for a in args:
            if a[0] == "kw":
                actual[a[1]] = a[2]
            else:
                actual[callee.args[pos]] = a
                pos += 1
        out = []
        for d in callee.args:
            ds = callee.syms[d]
            if d not in actual:
                if not ds.optional:
                    raise SyntaxError(f"{self.r.name}: call {name}: missing argument {d}")
                out.append("{}")
                continue
            a = actual[d]
            if ds.optional and not ds.rank and ds.base not in ("type", "procedure"):
                # an absent optional of the caller stays absent
                if a[0] == "name" and self.sym(a[1]) is not None and self.sym(a[1]).optional:
                    out.append(cname(a[1]))
                else:
                    out.append(f"Opt<{ds.ctype}>({self.ex(a)})")
            elif ds.base == "type" and a[0] == "name" and self.sym(a[1]) is not None and self.sym(a[1]).base == "type" and not self.sym(a[1]).dummy:
                out.append("*" + cname(a[1]))
            else:
                out.append(self.ex(a))
        return [self.line(f"{cname(name)}({', '.join(out)});")]


# --------------------------------------------------------------------------------------------- driver
TI_ROUTINES = [
    "atm_compute_vert_imp_coefs_work", "atm_compute_vert_imp_coefs",
    "atm_set_smlstep_pert_variables_work", "atm_set_smlstep_pert_variables",
    "atm_advance_acoustic_step_work", "atm_advance_acoustic_step",
    "atm_divergence_damping_3d",
    "atm_recover_large_step_variables_work", "atm_recover_large_step_variables",
    "atm_advance_scalars_work", "atm_advance_scalars",
    "atm_advance_scalars_mono_work", "atm_advance_scalars_mono",
    "atm_compute_dyn_tend_work", "atm_compute_dyn_tend",
    "atm_compute_solve_diagnostics_work", "atm_compute_solve_diagnostics",
    "atm_rk_integration_setup", "atm_compute_moist_coefficients", "atm_init_coupled_diagnostics",
    "atm_rk_dynamics_substep_finish",
    # regional (limited-area) path, config_apply_lbcs (TI:7198-7910)
    "atm_zero_gradient_w_bdy_work", "atm_zero_gradient_w_bdy", "atm_bdy_adjust_dynamics_speczone_tend",
    "atm_bdy_adjust_dynamics_relaxzone_tend", "atm_bdy_reset_speczone_values", "atm_bdy_adjust_scalars_work", "atm_bdy_adjust_scalars",
    "atm_bdy_set_scalars_work", "atm_bdy_set_scalars",
]


CORE_ROUTINES = ["atm_compute_mesh_scaling", "atm_compute_signs", "atm_compute_damping_coefs", "atm_adv_coef_compression",
                 "atm_couple_coef_3rd_order"]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    os.makedirs(out_dir, exist_ok=True)
    u = Unit()
    u.add_file(os.path.join(ref, "src/framework/mpas_constants.F"), ["CORE_ATMOSPHERE"], set())
    u.add_file(os.path.join(ref, "src/core_atmosphere/dynamics/mpas_atm_boundaries.F"), ["CORE_ATMOSPHERE"], set())
    u.add_file(os.path.join(ref, "src/core_atmosphere/mpas_atm_dimensions.F"), ["CORE_ATMOSPHERE"], set())
    u.add_file(os.path.join(ref, "src/core_atmosphere/dynamics/mpas_atm_time_integration.F"), ["CORE_ATMOSPHERE", "_MPI"], set(TI_ROUTINES))
    # init-time derivations of atm_mpas_init_block: pins mpas_model_b200/init_block.py
    u.add_file(os.path.join(ref, "src/core_atmosphere/mpas_atm_core.F"), ["CORE_ATMOSPHERE", "_MPI"], set(CORE_ROUTINES), module_spec=False)
    missing = [r for r in TI_ROUTINES + CORE_ROUTINES if r not in u.routines]
    if missing:
        raise SystemExit(f"f2cpp: routines not found in the reference: {missing}")
    body = u.emit_all(TI_ROUTINES + CORE_ROUTINES)
    with open(os.path.join(out_dir, "ti_ref.inc"), "w") as f:
        f.write("// GENERATED by oracle/f2cpp.py from the reference's Fortran source -- derived reference text, never commit.\n" + body)
    print(f"f2cpp: {len(TI_ROUTINES) + len(CORE_ROUTINES)} routines, {len(body.splitlines())} lines -> {out_dir}/ti_ref.inc")


if __name__ == "__main__":
    main()
