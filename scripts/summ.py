import json, sys
for line in sys.stdin:
    line = line.strip()
    if line.startswith("{") and line.endswith("}"):
        d = json.loads(line)
        print("value=%.4g ms/step=%.3f frac=%s n=%s phases=%s exch=%s" % (d["value"], d["ms_per_step"], d.get("roofline", {}).get("frac"), d.get("n_gpus"),
              d.get("phase_ms_rank0"), d.get("config", {}).get("exchange")))
    elif "dist]" in line or "Error" in line or "error" in line:
        print(line[:300])
