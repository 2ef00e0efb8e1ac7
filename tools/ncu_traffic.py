"""profiles/r01_traffic.json from an ncu CSV of `--metrics dram__bytes_read.sum,dram__bytes_write.sum` over one bench run:
mean DRAM bytes per launch of the conv / wgrad kernels (bench.py reads conv_tc2_kernel's figure for roofline.traffic)."""
import collections
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = collections.defaultdict(lambda: [0, 0.0])
ids = collections.defaultdict(set)
for r in rows:
    if r is hdr or not r[0].isdigit():
        continue
    name = r[ik]
    key = next((k for k in ("conv_tc2_kernel", "wgrad_tc_kernel") if k in name), None)
    if key is None or not r[im].startswith("dram__bytes"):
        continue
    per[key][1] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    ids[key].add(r[0])
out = {k: {"launches": len(ids[k]), "dram_bytes_per_launch": v[1] / max(1, len(ids[k]))} for k, v in per.items()}
out["source"] = f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the conv/wgrad launches of one bench run ({src})"
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
