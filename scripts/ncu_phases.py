"""Share of executed instructions and of warp-stall samples per source-line range of an ncu report.

    python scripts/ncu_phases.py <report.ncu-rep> file:lo-hi=name [...]   (unmatched lines -> their file name)
"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    rules = []
    for a in sys.argv[2:]:
        loc, name = a.split("=")
        f, rng = loc.split(":")
        lo, hi = rng.split("-")
        rules.append((f, int(lo), int(hi), name))
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    hdr, fname = None, ""
    inst, smp = collections.Counter(), collections.Counter()
    stall_cols = {}
    stalls = collections.defaultdict(collections.Counter)
    for r in csv.reader(txt.splitlines()):
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 6 and r[0] == "Line No":
            hdr = r
            i_s, i_x = hdr.index("# Samples"), hdr.index("Instructions Executed")
            stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith("stall_")}
            continue
        if hdr is None or len(r) != len(hdr) or r[0] == "":
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        g = fname
        for f, lo, hi, name in rules:
            if fname.startswith(f) and lo <= ln <= hi:
                g = name
                break
        inst[g] += int(r[i_x] or 0)
        smp[g] += int(r[i_s] or 0)
        for i, h in stall_cols.items():
            try:
                stalls[g][h] += int(r[i] or 0)
            except ValueError:
                pass
    ti, ts = sum(inst.values()) or 1, sum(smp.values()) or 1
    print("total: %d warp instructions, %d samples" % (ti, ts))
    for g, v in smp.most_common():
        top = ", ".join("%s %.0f%%" % (h[6:], 100. * n / max(1, sum(stalls[g].values())))
                        for h, n in stalls[g].most_common(4))
        print("%-14s samples %5.1f%%  inst %5.1f%%   %s" % (g, 100. * v / ts, 100. * inst[g] / ti, top))


if __name__ == "__main__":
    main()
