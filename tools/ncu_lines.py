"""Per CUDA source line: executed warp instructions, stall samples, average active lanes of one kernel in an
.ncu-rep captured with --import-source on (binary built with -lineinfo).
    python tools/ncu_lines.py rep.ncu-rep <kernel regex> [top N]"""
import csv
import io
import subprocess
import sys


def main(rep, pat, top=45):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + pat,
                          "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, agg = "", []
    hdr = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 8 and r[0].isdigit():
            ie, it, iss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
            try:
                agg.append((fname, int(r[0]), r[1].strip(), int(r[ie]), int(r[it]), int(r[iss])))
            except ValueError:
                pass
    tot = sum(a[3] for a in agg)
    ts = sum(a[5] for a in agg)
    print(f"{pat}: {tot / 1e6:.2f} M warp instructions, {ts} stall samples, {len(agg)} source lines")
    for a in sorted(agg, key=lambda a: -a[3])[:top]:
        print("%5.1f%% inst %5.1f%% stall  lanes %4.1f | %s:%d  %s" % (100 * a[3] / tot, 100 * a[5] / max(ts, 1), a[4] / max(a[3], 1), a[0][:14], a[1], a[2][:100]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 45)
