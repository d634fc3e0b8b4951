import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        if line: print(line[:200])
        continue
    d=json.loads(line)
    r=d["roofline"]
    print(d["config"]["workload"][:60], "| states/s %.3e ms/step %.3f e2e %.3e | gru frac %.3f (%.0f TF, %.1f us) whole frac %.3f | clocks %s %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["achieved"], r["us_per_launch"], r["whole_rollout"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
    print("   ", {k:round(v["us_per_launch"],1) for k,v in r["stages"].items()})
