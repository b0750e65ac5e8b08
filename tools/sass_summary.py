"""Per-kernel count of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma ->
UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG, setmaxnreg -> USETMAXREG, cluster barriers -> UCGABAR / UTCBAR)
in the shipped libfrx_b200.so.  Run in the build container (no GPU):  python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fancyrec_b200", "libfrx_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "USETMAXREG", "SYNCS", "HMMA",
        "FFMA", "STG", "LDG", "ATOM", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    stamp = open(os.path.join(ROOT, "fancyrec_b200", "libfrx_b200.stamp")).read().strip()
    print("library stamp (sha256 of sources + flags): %s" % stamp)
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k + (".2CTA" if ".2CTA" in op else "")] += 1
    demangle = subprocess.run(["c++filt"] + order, stdout=subprocess.PIPE, text=True).stdout.splitlines()
    print("%-88s %6s  %s" % ("kernel", "instr", "mnemonic counts"))
    for name, pretty in sorted(zip(order, demangle), key=lambda x: x[1]):
        c = counts[name]
        tags = "  ".join("%s %d" % (k, v) for k, v in sorted(c.items()) if k != "_total" and
                         not k.startswith(("FFMA", "STG", "LDG", "ATOM", "RED", "SYNCS")))
        short = re.sub(r"\(.*", "", pretty)[:88]
        print("%-88s %6d  %s" % (short, c["_total"], tags))


if __name__ == "__main__":
    main()
