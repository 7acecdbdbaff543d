import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench.json'))
print(f"value={d['value']:.1f} tiles/s  ms/step={d['ms_per_step']:.2f}  e2e={d['e2e']['value']:.1f}  launches={d['gpu_launches']}  clocks={d['clocks']}")
print(f"roofline: {d['roofline']['achieved']:.1f} TF = {d['roofline']['frac']:.3f}; model {d['model_tflops']:.1f} TF = {d['model_frac_of_tensor_peak']:.3f}; cpu={d['cpu_baseline']}")
for k, v in d['kernels'].items():
    print(f"  {k:14s}", {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
for k, v in d.get('gemm_modes', {}).items():
    print(f"    gemm {k:22s} {v['ms_per_step']:8.3f} ms  {v['tflops']:8.1f} TF")
