import json, sys
b=json.load(open('gpurun_out/bench.json'))
print('value %.1f fps  e2e %.1f  ms/step %.1f  roofline %.1f TF (%.1f%%)'%(b['value'],b['e2e']['value'],b['ms_per_step'],b['roofline']['achieved'],100*b['roofline']['frac']))
for k,v in b['kernels'].items():
    if k!='v2v': print('  %-16s %.4f ms/frame  %.0f GB/s'%(k,v['ms_per_frame'],v['GB_per_s']))
print('  v2v', b['kernels']['v2v'])
t=json.load(open('gpurun_out/v2v_ops.json'))
tot=sum(r['ms_per_frame'] for r in t)
print('total v2v ms/frame %.4f'%tot)
groups={}
for r in t:
    key=(r['kind'],r['cin'],r['cout'],r['k'],r['side'])
    g=groups.setdefault(key,[0,0.0,[]]); g[0]+=1; g[1]+=r['ms_per_frame']; g[2].append(r['tflops'] or 0)
for key,g in sorted(groups.items(), key=lambda kv:-kv[1][1]):
    print('  %-7s %3d->%3d k%d S%-3d x%-2d %.4f ms (%.1f%%)  %.0f TF'%(key[0],key[1],key[2],key[3],key[4],g[0],g[1],100*g[1]/tot,sum(g[2])/len(g[2])))
