import json, sys
d = json.load(open(sys.argv[1]))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1),
      "launches", d["gpu_launches"], "clocks", d["clocks"])
if d.get("roofline"):
    r = d["roofline"]; print("roofline", r["kernel"], round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3), "share", round(r["share_of_step"], 3))
tot = 0
for k, v in d["kernels"].items():
    tot += v["ms_per_step"]
    print(f"{k:34s} {v['launches']:3d} {v['ms_per_step']:8.3f} ms  {'' if v['tflops'] is None else round(v['tflops'],1)} {'' if v['gbs'] is None else round(v['gbs'],1)}")
print("sum of kernel ms", round(tot, 3))
if "cpu_baseline" in d: print(d["cpu_baseline"])
