import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        print(round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]) if "e2e" in d else None, d["phases_ms"], d["em_class_ms"], d["em_class_tasks"], d.get("em_pair_ms"), d.get("em_pair_tasks"))
