#!/usr/bin/env python
"""Print the per-kernel-class table of a bench.py JSON line:  python profiles/kern_table.py gpurun_out/x.log"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"ms/step {d['ms_per_step']:.2f}  fwd+bwd {d['value']/1e9:.3f} G/s ({d['roofline']['step_frac_fwdbwd']:.3f} of HBM roofline)  fwd {d['forward_only']['value']/1e9:.3f} G/s ({d['roofline']['step_frac_fwd']:.3f})  e2e {d['e2e']['value']/1e9:.3f} G/s")
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share"]):
    print(f"  {k:11s} share {v['share']:.3f}  avg {v['avg_ms']*1e3:7.1f} us x {v['launches_per_step']:5.1f}/step = {v['avg_ms']*v['launches_per_step']:.2f} ms" + (f"  {v['gbs']:.0f} GB/s" if 'gbs' in v else ""))
