import json,sys
for w in sys.argv[1:]:
    try:
        d=json.load(open(f"gpurun_out/bench_{w}.json")); r=d["roofline"]
        print(w, round(d["value"]), "evals/s", round(d["ms_per_step"],3), "ms/step frac", r["frac"] and round(r["frac"],3), "step_frac", round(r["step_frac"],3), d["config"]["chunks"], d["config"]["chunk_len"], d["config"]["halo"], "e2e", round(d["e2e"]["value"]), "lp_seam", round(d["e2e"]["log_prob_seam_evals_per_s"]), "cpu", d.get("cpu_baseline",{}).get("value"), {k:round(v,3) for k,v in r["kernel_share_of_step"].items()}, "launches", d["gpu_launches"])
    except Exception as e:
        print(w, "ERR", e)
