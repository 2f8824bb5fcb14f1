import json, sys
for f in sys.argv[1:]:
    l = [x for x in open(f) if x.startswith('{')]
    if not l: print(f, "no json"); continue
    d = json.loads(l[-1])
    print(f"== {f}: value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.1f} ({d['e2e'].get('ms_per_step',0):.3f} ms)  launches/step {d.get('gpu_launches',0)/max(d['steps'],1):.0f}  wall {d.get('wall_ms_per_step',0):.2f}")
    print("   steps:", d.get('ms_steps'))
    r = d.get('roofline'); print("   roofline:", r and {k: (round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k!='note'})
    tot = 0
    for k, v in (d.get('kernels') or {}).items():
        tot += v['ms_per_step']; print(f"   {k:28s} {v['launches_per_step']:6.1f}  {v['ms_per_step']:.4f} ms")
    print("   listed kernel total", round(tot,3), " cpu_baseline", d.get('cpu_baseline') and d['cpu_baseline']['value'])
