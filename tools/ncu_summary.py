"""Summarise ncu outputs brought back in gpurun_out/ (run in the build container, no GPU needed).

  python tools/ncu_summary.py launches <launches.csv>           # per-kernel time shares
  python tools/ncu_summary.py report <file.ncu-rep>             # key metrics of every captured launch
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpc__cycles_elapsed.avg.per_second",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = {}, 0.0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit.startswith("us") else v
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms
        a[1] += 1
        tot += ms
    print("%-90s %4s %10s %7s" % ("kernel", "n", "ms total", "share"))
    for k, (ms, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("%-90s %4d %10.3f %6.1f%%" % (k[:90], n, ms, 100 * ms / tot))
    print("%-90s %4s %10.3f" % ("TOTAL (serialised, cold-cache: compare shares)", "", tot))


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s  (grid %s x block %s)" % (d.get("Kernel Name", "?")[:100], d.get("launch__grid_size"), d.get("launch__block_size")))
        for k in KEYS:
            for h, u in zip(hdr, units):
                if h == k or h.endswith(k):
                    print("   %-80s %s %s" % (h, d[h], u))
                    break


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
