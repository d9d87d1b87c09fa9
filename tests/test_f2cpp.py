"""Known-answer tests of the Fortran -> C++ transliterator's expression grammar (oracle/f2cpp.py).

The oracle is pinned to the reference's own source THROUGH this transliterator (tests/test_reference_pin.py), so the one thing
it must get right on its own is how Fortran groups an expression: Fortran 2003 R722 -- `**` binds tightest and associates to the
RIGHT, then `*` `/` (left), then unary and binary `+` `-` (left; a leading sign applies to the whole first term), relational,
`.not.`, `.and.`, `.or.`, `.eqv.`.  The answers below are those rules applied by hand; the evaluator walks the parser's tree
without re-ordering anything, so a wrong grouping gives a wrong number."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import f2cpp  # noqa: E402


def _num(text):
    t = text.lower().split("_")[0].replace("d", "e")
    return float(t) if any(c in t for c in ".e") else int(t)


def _ev(n, env):
    k = n[0]
    if k == "num":
        return _num(n[1])
    if k == "log":
        return n[1]
    if k == "name":
        return env[n[1]]
    if k == "paren":
        return _ev(n[1], env)
    if k == "un":
        v = _ev(n[2], env)
        return {"-": lambda: -v, "+": lambda: +v, "!": lambda: not v}[n[1]]()
    if k == "bin":
        a, b = _ev(n[2], env), _ev(n[3], env)
        op = n[1]
        if op == "/" and isinstance(a, int) and isinstance(b, int):
            return int(a / b)                         # Fortran integer division truncates toward zero
        return {"+": lambda: a + b, "-": lambda: a - b, "*": lambda: a * b, "/": lambda: a / b, "**": lambda: a ** b,
                "==": lambda: a == b, "/=": lambda: a != b, "<": lambda: a < b, "<=": lambda: a <= b, ">": lambda: a > b,
                ">=": lambda: a >= b, "&&": lambda: a and b, "||": lambda: a or b}[op]()
    raise AssertionError(n)


CASES = [
    ("2.0 - 3.0 - 4.0", {}, -5.0),                       # left associative
    ("2.0 ** 3.0 ** 2.0", {}, 512.0),                    # right associative: 2 ** (3 ** 2)
    ("-2.0 ** 2.0", {}, -4.0),                           # the sign applies to the power
    ("-2.0 * 3.0 + 1.0", {}, -5.0),
    ("8.0 / 4.0 * 2.0", {}, 4.0),                        # (8 / 4) * 2, not 8 / (4 * 2)
    ("2.0 * 3.0 ** 2.0", {}, 18.0),
    ("2.0 ** -1.0", {}, 0.5),                            # signed exponent
    ("1.0 + 2.0 * 3.0 - 4.0 / 2.0", {}, 5.0),
    ("7 / 2 * 2", {}, 6),                                # integer division first
    ("1.0_RKIND + 2.5e-1 - 1.0d0", {}, 0.25),            # kind suffix, e and d exponents
    ("a - (b - c)", dict(a=1.0, b=2.0, c=3.0), 2.0),     # parentheses are kept
    ("x > 1.0 .and. y <= 2.0 .or. z == 3", dict(x=0.0, y=0.0, z=3), True),       # (x > 1 and y <= 2) or z == 3
    ("a .or. b .and. c", dict(a=True, b=False, c=False), True),                  # a or (b and c)
    ("a .and. .not. b .or. c", dict(a=True, b=True, c=False), False),            # (a and (not b)) or c
    (".not. a .and. b", dict(a=False, b=False), False),                          # (not a) and b
    # mpas_atm_time_integration.F:4479, as written there: the second test is NOT guarded by config_apply_lbcs
    ("config_apply_lbcs .and. (m == nRelaxZone) .or. (m == nRelaxZone-1)", dict(config_apply_lbcs=False, m=4, nrelaxzone=5), True),
    # mpas_atm_core.F:1396: the sign applies to the whole term ((dc ** 2) * a) / 12
    ("- (dc **2) * a / 12.", dict(dc=3.0, a=4.0), -3.0),
]


@pytest.mark.parametrize("text,env,want", CASES, ids=[c[0] for c in CASES])
def test_expression_grouping(text, env, want):
    got = _ev(f2cpp.parse_expr(text), env)
    assert got == want and type(got) is type(want)


def test_references_and_sections_parse():
    """Array references, derived-type members, keyword arguments and sections keep their structure."""
    e = f2cpp.parse_expr("zb(:,1,edgesOnCell(i,iCell))")
    assert e[0] == "call" and e[1] == ("name", "zb") and e[2][0] == ("section", None, None) and e[2][2][0] == "call"
    e = f2cpp.parse_expr("block % configs")
    assert e == ("member", ("name", "block"), "configs")
    e = f2cpp.parse_expr("max(a, b, dim=1)")
    assert e[2][2] == ("kw", "dim", ("num", "1"))
    with pytest.raises(SyntaxError):
        f2cpp.parse_expr("a + * b")
