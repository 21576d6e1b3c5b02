import sys, json
for path in sys.argv[1:]:
    l=[x for x in open(path) if x.startswith("{")]
    if not l: print(path, "no json"); continue
    d=json.loads(l[-1]); e=d["e2e"]
    q=e.get("queue64") or {}
    print(path.split("/")[-1], "value", round(d["value"],1), "e2e", round(e["value"],1), "ms", round(e["ms_per_step"],3), "18pl", round(e["all_18_planes"]["value"],1), round(e["all_18_planes"]["ms_per_step"],2), "queue64", round(q.get("value",0),1), "d2h", e["d2h_bytes_per_step"], "raster_in", round(e["raster_in"]["value"],1))
