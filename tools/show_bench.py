import json, sys
d = json.load(open(sys.argv[1]))
print(round(d["value"] / 1e6, 1), "Mpos/s", round(d["ms_per_step"], 2), "ms/step  inflate out", round(d["inflate_out_gbs"], 1), "GB/s",
      {k: round(v, 2) for k, v in d["roofline"]["stage_ms"].items()}, d.get("inflate_counters"))
