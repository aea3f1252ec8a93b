#!/usr/bin/env python
"""Development tool: one-screen summary of a bench.py JSON line."""
import json
import sys

for path in sys.argv[1:]:
    d = json.load(open(path))
    par = d.get("parity") or {}
    print(f"== {path}: N={d['n_gpus']} value {d['value']:.4e} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['value']:.3e} frac {d['roofline']['frac']:.3f} "
          f"kernel_ms {['%.4f' % k for k in d['roofline']['kernel_ms_per_rank']]} parity {par.get('rho_rel_linf_vs_cpu_reference')} "
          f"replicas {(par.get('replicas') or {}).get('bit_identical')} cpu {(d.get('cpu_baseline') or {}).get('value')}")
    for e in d.get("other_workloads", []):
        if "error" in e:
            print("   ", e)
            continue
        p = e.get("parity") or {}
        print(f"    {e['workload'][:44]:44s} {str(e.get('scaling')):6s} pps {e['point_steps_per_s']:.3e} ms {e['ms_per_step']:9.3f} frac {e['fp64_frac']:.3f} "
              f"parity {p.get('rho_rel_linf_vs_cpu_reference')} ok {p.get('ok')} repl {(p.get('replicas') or {}).get('bit_identical')} x_refcuda {e.get('speedup_vs_reference_cuda_one_gpu')}")
    for f in d.get("full_run", []):
        print("    full_run", {k: f[k] for k in f if k in ("workload", "wall_s", "mean_s_per_time_step", "energy_trace_rel_err", "reference_cpu_wall_s", "error")})
