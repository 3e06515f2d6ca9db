import sys,re
reg=eval(sys.argv[1])
fname=sys.argv[2] if len(sys.argv)>2 else 'pd_stage_b_mma.cuh'
acc={r[0]:[0,0] for r in reg}; other=[0,0]
for ln in sys.stdin:
    m=re.match(r'\s*([\d.]+)% inst\s+([\d.]+)% stall\s+(\S+):(\d+)',ln)
    if not m:
        if ln.startswith('total'): print(ln.strip())
        continue
    i,s,f,l=float(m[1]),float(m[2]),m[3],int(m[4])
    if f!=fname:
        other[0]+=i; other[1]+=s
        if s>1.5: print('other',ln.strip()[:110])
        continue
    for n,a,b in reg:
        if a<=l<b: acc[n][0]+=i; acc[n][1]+=s
for k,v in acc.items(): print(f'{k:16s} inst {v[0]:5.1f}%  stall {v[1]:5.1f}%')
print('other',other)
