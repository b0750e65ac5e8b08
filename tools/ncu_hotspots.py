"""Top SASS instructions of one kernel of an .ncu-rep by warp-stall samples (runs in the build container).

  python tools/ncu_hotspots.py <file.ncu-rep> <kernel-name regex> [launch index among the matches] [top N]
"""
import csv
import io
import subprocess
import sys

path, regex = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + regex,
                      "--launch-skip", str(skip), "--launch-count", "1"], stdout=subprocess.PIPE, text=True).stdout
lines = out.splitlines()
name = lines[0]
rows = list(csv.DictReader(io.StringIO("\n".join(lines[1:]))))
def num(v):
    try:
        return int(v)
    except (TypeError, ValueError):
        return -1
rows = [r for r in rows if num(r.get("# Samples")) >= 0 and (r.get("Address") or "").startswith("0x")]   # SASS view only
seen, uniq = set(), []
for r in rows:                                     # the export lists every SASS line once per view
    if r["Address"] not in seen:
        seen.add(r["Address"])
        uniq.append(r)
rows = uniq
tot = sum(num(r["# Samples"]) for r in rows)
ins = sum(max(num(r["Instructions Executed"]), 0) for r in rows)
print(name[:160])
print("warp-stall samples: %d ; warp instructions executed: %d ; SASS lines: %d" % (tot, ins, len(rows)))
print("%7s %7s  %12s  %s" % ("samples", "share", "warp instr", "SASS"))
for r in sorted(rows, key=lambda r: -num(r["# Samples"]))[:top]:
    n = num(r["# Samples"])
    print("%7d %6.1f%%  %12s  %s" % (n, 100.0 * n / max(tot, 1), r["Instructions Executed"], r["Source"].strip()[:110]))
