import json, sys
d = json.load(open(sys.argv[1]))
r = d["roofline"]
print(f"value={d['value']:.1f} img/s  ms/step={d['ms_per_step']:.2f}  e2e={d['e2e']['value'] if d.get('e2e') else None}  launches={d['gpu_launches']} clocks={d['clocks']}")
print(f"gemm: {r['achieved']:.0f} TF/s frac={r['frac']:.3f} (burst {r['frac_of_burst_peak']:.3f}) share={r['share_of_step']:.3f} whole-step {r['whole_step_tflops']:.0f} TF/s frac {r['whole_step_frac']:.3f}")
print("head:", d.get("roofline_head"))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print(f"  {k:14s} {v['ms_per_step']:8.3f} ms  x{v['launches_per_step']:3d}  share {v['share']:.3f}" + (f"  {v['tflops']:.0f} TF/s" if 'tflops' in v else ""))
if d.get("profiled_step"): print("profiled step:", d["profiled_step"])
print("sum kernels ms:", sum(v["ms_per_step"] for v in d["kernels"].values()))
if d.get("cpu_baseline"): print("cpu:", d["cpu_baseline"])
