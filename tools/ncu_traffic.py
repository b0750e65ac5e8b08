"""profiles/score_topk_traffic.json from an `ncu --set full` report: DRAM bytes per launch of the fused score + top-k
kernel (bench.py's roofline.traffic), stamped with the library build it was captured from.
    python tools/ncu_traffic.py gpurun_out/<report>.ncu-rep
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fancyrec_b200 import _lib  # noqa: E402

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def gb(row, key):
    v, u = float(row[ix[key]]), units[ix[key]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]


vals = []
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    if "score_kernel<0" in name.replace(" ", "") or "score_kernel<(int)0" in name.replace(" ", ""):
        ms = float(r[ix["gpu__time_duration.sum"]])
        ms = ms if units[ix["gpu__time_duration.sum"]] in ("ms", "msecond") else ms / 1e3
        if ms > 1.0:                    # the main launch, not the strided sample pass
            vals.append((gb(r, "dram__bytes_read.sum"), gb(r, "dram__bytes_write.sum"), ms))
assert vals, "no fused top-k launch in the report"
rd = sum(v[0] for v in vals) / len(vals)
wr = sum(v[1] for v in vals) / len(vals)
doc = {"kernel": "frx::score_kernel<MODE_TOPK, bf16>", "launches_averaged": len(vals),
       "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr,
       "ncu_duration_ms": sum(v[2] for v in vals) / len(vals), "library_stamp": _lib.source_hash(),
       "source": "ncu --set full capture %s of `bench.py --steps 2` (config 2)" % os.path.basename(rep)}
json.dump(doc, open(os.path.join(ROOT, "profiles", "score_topk_traffic.json"), "w"), indent=1)
print(json.dumps(doc, indent=1))
