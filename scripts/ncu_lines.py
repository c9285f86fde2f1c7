"""Per-CUDA-line warp-stall samples and executed instructions from an ncu report (--import-source on).

    python scripts/ncu_lines.py <report.ncu-rep> [top]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = None
    fname = ""
    agg = []
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 6 and r[0] == "Line No":
            hdr = r
            i_s, i_x = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[0] != "":  # a CUDA line row (aggregated over its SASS)
            try:
                agg.append((fname, int(r[0]), r[1].strip(), int(r[i_s] or 0), int(r[i_x] or 0)))
            except ValueError:
                pass
    tot_s = sum(a[3] for a in agg) or 1
    tot_x = sum(a[4] for a in agg) or 1
    print("total samples %d, warp instructions %d" % (tot_s, tot_x))
    for f, ln, src, s, x in sorted(agg, key=lambda a: -a[3])[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100. * s / tot_s, 100. * x / tot_x, f, ln, src[:100]))


if __name__ == "__main__":
    main()
