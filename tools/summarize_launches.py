"""Turn an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list into the
per-kernel summary and the tap-GEMM traffic record committed under profiles/ (bench.py reads the latter)."""
import collections
import csv
import json
import re
import sys

src, out_txt, out_json = sys.argv[1:4]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
L = collections.OrderedDict()
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}
for r in rows[hi + 1:]:
    if len(r) < 15:
        continue
    d = L.setdefault(r[0], {"name": r[4], "stream": r[6]})
    v = float(r[14].replace(",", ""))
    if r[12] == "gpu__time_duration.sum":
        d["us"] = v * TIME[r[13]]
    elif r[12] == "dram__bytes_read.sum":
        d["rd"] = v * BYTES[r[13]]
    elif r[12] == "dram__bytes_write.sum":
        d["wr"] = v * BYTES[r[13]]


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("at::native::", "native::")
    return n[:70]


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in L.values():
    a = agg[short(d["name"])]
    a[0] += 1
    a[1] += d["us"]
    a[2] += d.get("rd", 0)
    a[3] += d.get("wr", 0)
tot = sum(a[1] for a in agg.values())
out = ["# one eager training micro-step (step 2 of an accumulation cycle) of the CS UNet under ncu",
       "# command: ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
       "--clock-control none --csv python tools/ncu_step.py",
       f"# total {tot / 1e3:.2f} ms over {len(L)} launches (cold-cache, serialised: compare SHARES, not absolutes); "
       f"streams seen: {sorted(set(d['stream'] for d in L.values()))}",
       "share%   ms   launches  dram_rd_MB/launch  dram_wr_MB/launch  kernel"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    out.append(f"{100 * a[1] / tot:6.2f} {a[1] / 1e3:8.3f} {a[0]:6d} {a[2] / a[0] / 1e6:12.2f} {a[3] / a[0] / 1e6:12.2f}  {k}")
for key in ("tapconv_kernel", "wgrad_kernel"):
    sel = [d for d in L.values() if key in d["name"]]
    t = sum(d["us"] for d in sel)
    out.append(f"# {key} (all instantiations): {len(sel)} launches, {t / 1e3:.3f} ms = {100 * t / tot:.1f}% of the step; "
               f"DRAM traffic per launch {sum(d.get('rd', 0) + d.get('wr', 0) for d in sel) / len(sel) / 1e6:.2f} MB")
open(out_txt, "w").write("\n".join(out) + "\n")
tc = [d for d in L.values() if "tapconv_kernel" in d["name"]]
json.dump({"kernel": "tapconv_kernel", "launches_per_step": len(tc),
           "dram_bytes_per_launch": sum(d.get("rd", 0) + d.get("wr", 0) for d in tc) / len(tc),
           "share_of_step": sum(d["us"] for d in tc) / tot, "source": src}, open(out_json, "w"), indent=1)
print("\n".join(out[:16] + out[-2:]))
