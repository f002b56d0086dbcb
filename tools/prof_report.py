import json,sys
from collections import defaultdict
d=json.load(open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/bench_profile.json'))
fam=d['families']; tot=sum(f['ms'] for f in fam.values())
for k,f in sorted(fam.items(), key=lambda kv:-kv[1]['ms'])[:12]:
    tf = f['flops']/(f['ms']*1e-3)/1e12 if f['flops'] else 0
    gb = f['bytes']/(f['ms']*1e-3)/1e9 if f['bytes'] else 0
    print(f"{k:28s} n={f['n']:4d} ms={f['ms']:9.3f} share={f['ms']/tot:6.3f} TF/s={tf:8.1f} GB/s={gb:8.1f}")
print('total ms', tot)
for nm in ('ns_gemm_nt','ns_gemm_tn','ns_attention_fwd','ns_attention_bwd'):
    g=defaultdict(lambda:[0,0.0])
    for r in d['launches']:
        if r['name']==nm:
            key=round(r['flops']/1e9,1); g[key][0]+=1; g[key][1]+=r['ms']
    for k,(n,ms) in sorted(g.items(), key=lambda kv:-kv[1][1])[:9]:
        print(f"  {nm[3:]:14s} GF={k:8.1f} n={n:3d} total ms={ms:7.3f} avg={ms/n:6.3f} TF/s={k*n/ms:7.1f}")
