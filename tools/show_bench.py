import json, sys
d = json.load(open(sys.argv[1]))
print("value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1),
      "host_ms/batch", d["e2e"].get("host_ms_per_batch"), "launches", d["gpu_launches"], "graph", d.get("cuda_graph"), "clocks", d["clocks"])
if d.get("loss_check"): print("loss_check", d["loss_check"])
if d.get("roofline"):
    r = d["roofline"]; print("roofline", r["kernel"], round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3), "share", round(r["share_of_step"], 3))
tot = 0
for k, v in d.get("kernels", {}).items():
    tot += v["ms_per_step"]
    print(f"{k:34s} {v['launches']:3d} {v['ms_per_step']:8.3f} ms  {'' if v['tflops'] is None else round(v['tflops'],1)} {'' if v['gbs'] is None else round(v['gbs'],1)} {'' if not v.get('frac_hbm') else round(v['frac_hbm'],3)}")
print("sum of kernel ms", round(tot, 3))
for k in ("setconv_encoder", "multivar", "grid608", "inference", "cpu_baseline"):
    if k in d: print(k, {kk: vv for kk, vv in d[k].items() if kk not in ("roofline",)})
