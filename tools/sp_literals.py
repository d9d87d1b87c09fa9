#!/usr/bin/env python
"""PRECISION=single build: rewrite every un-suffixed floating literal of the CUDA sources as a float literal.

The reference's single build makes default-kind literals single precision (reference Makefile:861-873,
src/framework/mpas_kind_types.F:22-28), so `0.5 * x` stays a float operation there.  C++ literals are double unless
suffixed and nvcc has no -fsingle-precision-constant, so the sources of libmpasb_sp.so are passed through this filter
first (csrc/Makefile): comments, string and character literals are left alone; pow_cr.cuh (double-double arithmetic on
purpose) is not filtered.  usage: sp_literals.py <src dir> <dst dir>"""
import os
import re
import sys

LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")
SKIP = re.compile(r'//[^\n]*|/\*.*?\*/|"(?:\\.|[^"\\])*"|\'(?:\\.|[^\'\\])*\'', re.S)


def convert(text):
    out, pos = [], 0
    for m in SKIP.finditer(text):
        out.append(LIT.sub(lambda t: t.group(1) + "f", text[pos:m.start()]))
        out.append(m.group(0))
        pos = m.end()
    out.append(LIT.sub(lambda t: t.group(1) + "f", text[pos:]))
    return "".join(out)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        if not name.endswith((".cu", ".cuh", ".inl")) or name.startswith("pow_cr"):
            continue
        with open(os.path.join(src, name)) as f:
            text = convert(f.read())
        path = os.path.join(dst, name)
        if not os.path.exists(path) or open(path).read() != text:
            with open(path, "w") as f:
                f.write(text)


if __name__ == "__main__":
    main()
