import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    j = json.loads(line)
    r = j.get("roofline") or {}
    e = j.get("e2e") or {}
    c = j.get("cpu_baseline") or {}
    print("value %.4g %s | ms/step %.3f | mover %.3f ms import %.3f ms | roofline frac %.3f (whole step %.3f) | e2e %.3g | cpu %.3g (%s cores) | clocks %s | gpus %s deposit %s" % (
        j["value"], j["unit"], j["ms_per_step"], r.get("launch_ms", 0), r.get("import_ms", 0), r.get("frac", 0), r.get("whole_step_frac", 0),
        e.get("value", 0), c.get("value", 0), c.get("cores"), j.get("clocks"), j.get("n_gpus"), j.get("config", {}).get("deposit")))
