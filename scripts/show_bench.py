"""Pretty-print gpurun_out/bench.log (the JSON line) and gpurun_out/bench_kernels.log."""
import json, sys
def show_bench(path):
    d=json.loads(open(path).read().strip().splitlines()[-1])
    k=d.pop('kernels',{})
    print(f"value {d['value']} {d['unit']}  ms/step {d['ms_per_step']}  step_tflops {d.get('step_tflops')} frac {d.get('step_tensor_frac_sustained')}  e2e {d.get('e2e',{}).get('value') if d.get('e2e') else None}  clocks {d.get('clocks')}")
    print("roofline", d.get('roofline')); print("cpu", d.get('cpu_baseline'))
    for n,e in sorted(k.items(), key=lambda x:-x[1]['ms_per_step']): print(f"  {n:24s} {e['ms_per_step']:8.3f} ms  share {e.get('share', e.get('share_of_kernel_time')):.3f}  {e.get('achieved')} {e.get('unit')} frac {e.get('frac')}")
    print("kernel share of step", d.get('kernel_time_share_of_step'))
    for sub in ('train','long','v2','gpu_baseline'):
        r=d.get(sub)
        if not r: continue
        r=dict(r); kk=r.pop('kernels',None); r.pop('config',None)
        print(sub.upper(), json.dumps(r)[:1500])
        if kk:
            for n,e in sorted(kk.items(), key=lambda x:-x[1]['ms_per_step']): print(f"    {n:24s} {e['ms_per_step']:8.3f} ms x{e['launches_per_step']}  {e.get('us_per_launch','')}")
def show_kernels(path):
    for l in open(path):
        try: d=json.loads(l)
        except Exception: print(l.strip()); continue
        print(f"  {d['case']:28s} cg={d.get('cta_pair','-')} bn={d.get('block_n','-')} ms={d['ms_median']:.4f} tflops={d.get('tflops','-')} frac={d.get('frac_burst','-')} gbs={d.get('gbs','-')} frac_hbm={d.get('frac_hbm','-')}")
if __name__=='__main__':
    show_bench(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/bench.log')
    if len(sys.argv)<=1: show_kernels('gpurun_out/bench_kernels.log')
