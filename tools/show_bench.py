"""prints the key numbers of bench.py JSON lines (skipping the other lines some libraries print to stdout)"""
import json, sys
for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        out = {k: d.get(k) for k in ("n_gpus", "value", "ms_per_step")}
        out["e2e"] = d.get("e2e", {}).get("value") if d.get("e2e") else None
        out["stage_ms"] = {k: round(v, 2) for k, v in (d.get("stage_ms") or {}).items()}
        if d.get("parity"):
            out["parity_ok"] = d["parity"]["ok"]
        if d.get("pcg_profile"):
            out["pcg_us_it"] = {k: round(v, 1) for k, v in d["pcg_profile"]["us_per_iteration"].items()}
            out["pcg_setup_us"] = round(d["pcg_profile"]["setup_us_per_solve"], 0)
            out["pcg_per_rank"] = d["pcg_profile"].get("per_rank")
        if d.get("per_step"):
            out["events"] = d["per_step"].get("events")
        print(path, json.dumps(out))
