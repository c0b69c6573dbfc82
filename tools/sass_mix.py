#!/usr/bin/env python
"""Static instruction mix of a kernel and of its innermost backward-branch loops.

usage: python tools/sass_mix.py OBJECT.o MANGLED_KERNEL_NAME
Prints the opcode histogram of the whole kernel and of every loop (a backward branch and its target), e.g. the rolled line loop of
phase B of ns_vmult_r1_kernel, whose body runs three times per decade."""
import collections
import re
import subprocess
import sys

obj, fun = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True, check=True).stdout
ins = []
for line in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)(\.[\w.]+)?\s*(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2), m.group(4)))


def mix(items):
    c = collections.Counter(op for _, op, _ in items)
    fp64 = c["DFMA"] + c["DMUL"] + c["DADD"]
    top = ", ".join(f"{op} {n}" for op, n in c.most_common(14))
    return f"{len(items)} instructions, fp64 {fp64} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']}); {top}"


print("kernel:", mix(ins))
for addr, op, rest in ins:
    if op == "BRA":
        m = re.search(r"0x([0-9a-f]+)", rest)
        if m and int(m.group(1), 16) < addr:
            lo = int(m.group(1), 16)
            body = [i for i in ins if lo <= i[0] <= addr]
            if len(body) > 50:
                print(f"loop 0x{lo:x}..0x{addr:x}:", mix(body))
