#!/usr/bin/env python
"""SASS instruction budget of the hot kernels: for every backward branch (= loop) of a kernel, the number of SASS
instructions between the branch target and the branch, plus an opcode histogram of the innermost hot loop.

    python tools/sass_budget.py [substring of the mangled kernel name ...]

The composite kernels are FP32-issue bound, so instructions per blended (quad, splat) pair are the
quantity to optimise; this prints them from the built library (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gaussian-mesh-splatting_b200", "gms_b200", "libgms_b200.so")
INSN = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")


def kernels(patterns):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name and any(p in name for p in patterns):
                yield name, body
            name, body = m.group(1), []
            continue
        m = INSN.match(line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name and any(p in name for p in patterns):
        yield name, body


def main():
    pats = sys.argv[1:] or ["k_composite_bwd5ILi6ELb0", "k_composite_bwd3ILi6ELb0", "k_composite_fwd2ILb1"]
    for name, body in kernels(pats):
        print(f"== {name}: {len(body)} SASS instructions")
        addr_index = {a: i for i, (a, _) in enumerate(body)}
        loops = []
        for i, (a, txt) in enumerate(body):
            m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", txt)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a and tgt in addr_index:
                    loops.append((addr_index[tgt], i))
        for lo, hi in sorted(loops, key=lambda t: t[1] - t[0]):
            ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body[lo:hi + 1])
            top = ", ".join(f"{k} {v}" for k, v in ops.most_common(8))
            print(f"   loop 0x{body[lo][0]:04x}..0x{body[hi][0]:04x}: {hi - lo + 1:4d} instructions   [{top}]")


if __name__ == "__main__":
    main()
