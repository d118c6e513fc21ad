"""Pretty-print the JSON line of a bench.py run: python scripts/show_bench.py <file>"""
import json
import sys

for l in open(sys.argv[1]):
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    print("%s  value %.1f %s  ms/step %.3f  n_gpus %d" % (d.get("impl", "b200"), d["value"], d["unit"], d["ms_per_step"], d["n_gpus"]))
    if d.get("roofline"):
        r = d["roofline"]
        print("roofline: %s %.0f GB/s frac %.3f | step %.0f GB/s frac %.3f" % (r["kernel"], r["achieved"], r["frac"], r["step_achieved"], r["step_frac"]))
    for k in d.get("kernels", []):
        print("  %-20s %7.3f ms/step  n=%.0f  %s" % (k["kernel"], k["ms_per_step"], k["launches_per_step"],
                                                 ("%.0f GB/s (%.2f)" % (k["achieved_GBs"], k["frac"])) if k.get("achieved_GBs") else ""))
    print("poisson ms", d.get("poisson_solve_ms"), "e2e", d.get("e2e") and d["e2e"]["value"], "clocks", d.get("clocks"), "check", d.get("check"))
