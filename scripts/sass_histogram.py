"""Opcode histogram of one kernel of libqgsb.so (cuobjdump -sass), the instruction-mix evidence quoted in DESIGN.md.

    python scripts/sass_histogram.py <kernel-name-substring> [<second substring> ...] [--lib path]

Prints, for the FIRST function whose (mangled) name contains all substrings: instruction count, code bytes, and the
counts per opcode (modifiers stripped) in decreasing order.
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    lib = os.path.join(REPO, "qgs_b200", "libqgsb.so")
    if "--lib" in sys.argv:
        lib = sys.argv[sys.argv.index("--lib") + 1]
        args = [a for a in args if a != lib]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name, ops, inside = None, collections.Counter(), False
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if inside:
                break
            inside = all(a in m.group(1) for a in args)
            name = m.group(1) if inside else name
            continue
        if not inside:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            ops[m.group(1)] += 1
    if not ops:
        sys.exit("no function matches %s" % (args,))
    total = sum(ops.values())
    print("function     %s" % name)
    print("instructions %d  (%d bytes of code)" % (total, 16 * total))
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print("FP64         %d  (%.1f %%)" % (fp64, 100. * fp64 / total))
    for op, n in ops.most_common():
        print("%-12s %6d  %5.1f %%" % (op, n, 100. * n / total))


if __name__ == "__main__":
    main()
