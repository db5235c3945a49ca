"""Print the essentials of bench JSON lines: python tools/show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(path, round(d["value"]), "evals/s", round(d["ms_per_step"], 3), "ms/step | frac", r["frac"] and round(r["frac"], 3),
              "step_frac", round(r["step_frac"], 3), "| chunks", d["config"]["chunks"], d["config"]["chunk_len"], d["config"]["halo"],
              "| e2e", round(d["e2e"]["value"]), "| cpu", d.get("cpu_baseline", {}).get("value"),
              "| us", {k: round(v, 1) for k, v in r["kernel_us_per_launch"].items()}, "| launches", d["gpu_launches"],
              "| frozen", r.get("frozen_frac"), "| parity", d.get("parity_check"))
    except Exception as e:
        print(path, "ERR", e)
