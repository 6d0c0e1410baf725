import json,sys
for ln in open(sys.argv[1]):
    d=json.loads(ln); l=d["line"]
    if l.get("failed"): print(d["variant"], d["workload"], "FAILED", l.get("stderr","")[-300:]); continue
    k=l["kernels_ms_per_build"]; p=l.get("parity",{})
    print("%-34s %-10s total %.3f e2e %.3f | dens %.3f vmat %.3f basis %.3f formg %.3f func %.3f fin %.3f | dE %.1e dV %.1e" % (d["variant"], d["workload"], l["ms_per_step"], (l.get("e2e") or {}).get("ms_per_step",0), k["k_density"],k["k_scatter"],k["k_basis"],k["k_form_g"],k["k_functional"],k["finish"], p.get("dE_xc",-1), p.get("max_dV_xc",-1)))
