#!/usr/bin/env python
"""Prints one line per bench JSON file (scratch helper for gpurun experiments)."""
import glob, json, sys
for pat in sys.argv[1:]:
    for f in sorted(glob.glob(pat)):
        try:
            d = json.loads(open(f).read().strip().splitlines()[-1])
        except Exception as ex:
            print(f, "ERR", ex); continue
        r = d.get("roofline", {}); lm = r.get("layer_ms", {})
        e2e = d.get("e2e", {}).get("value", 0)
        print(f"{f.split('/')[-1][:-5]:30s} sims/s {d['value']/1e6:6.3f}M e2e {e2e/1e6:5.2f}M ms/step {d['ms_per_step']:7.2f} evals/s {d.get('net_evals_per_s',0)/1e6:5.2f}M "
              f"clk {d['clocks']['sm_mhz']} pw {d['clocks'].get('power_w_max')} | " + " ".join(f"{k}={v:.3f}" for k, v in lm.items()))
