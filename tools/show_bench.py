import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value %.0f evals/s  ms/step %.3f  wall ms/step %.3f' % (d['value'], d['ms_per_step'], d.get('wall_ms_per_step', 0)))
print('e2e states %.0f  planes %.0f' % (d['e2e']['value'], d['e2e']['planes']['value']))
print('clocks', d['clocks'], 'cpu', d.get('cpu_baseline'))
r = d['roofline']; print('roofline %.1f TF frac %.3f (%s) whole-net %.1f TF' % (r['achieved'], r['frac'], r['kernel'], r.get('whole_net_tflops_per_gpu', r.get('whole_net_tflops', 0.0))))
for k in d['kernels']:
    print(f"{k['kernel']:22s} x{k['launches_per_step']:2d} {k['us_per_launch']:8.1f} us  share {k['share']:.3f}  {k['tflops'] and round(k['tflops'],1)} TF  {k['gbs']:.0f} GB/s")
