import csv,re,collections,subprocess,sys
subprocess.run("ncu -i "+(sys.argv[2] if len(sys.argv)>2 else "gpurun_out/prof_fl.ncu-rep")+" --page source --csv --kernel-name regex:logmel > /tmp/fl_src.csv 2>/dev/null", shell=True)
rows=list(csv.reader(open('/tmp/fl_src.csv')))
hidx=[i for i,r in enumerate(rows) if r and r[0]=='Address']
h=rows[hidx[0]]; col={n:i for i,n in enumerate(h)}
ins=rows[hidx[0]+1:(hidx[1]-1 if len(hidx)>1 else len(rows))]
def f(r,n):
    try: return float(r[col[n]])
    except: return 0.0
stalls=[n for n in h if n.startswith('stall_') and 'Not' not in n]
tot=sum(f(r,'# Samples') for r in ins)
def opc(r):
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[col['Source']]); return m.group(2).split('.')[0] if m else '?'
allst=collections.Counter()
for r in ins:
    for n in stalls: allst[n[6:]]+=f(r,n)
print('total instr/frame', sum(f(r,'Instructions Executed') for r in ins)/192064)
print('stalls %:',' '.join(f"{k}:{100*v/tot:.1f}" for k,v in allst.most_common(10)))
chunk=int(sys.argv[1]) if len(sys.argv)>1 else 80
for s in range(0,len(ins),chunk):
    sub=ins[s:s+chunk]
    smp=sum(f(r,'# Samples') for r in sub)
    if smp/tot<0.012: continue
    st=collections.Counter()
    for r in sub:
        for n in stalls: st[n[6:]]+=f(r,n)
    byop=collections.Counter()
    for r in sub: byop[opc(r)]+=f(r,'# Samples')
    ex=sum(f(r,'Instructions Executed') for r in sub)/192064
    print(f"{s:5d}: {100*smp/tot:5.1f}% exec/frame {ex:5.1f} | "+' '.join(f"{k}:{100*v/tot:.1f}" for k,v in st.most_common(4))+' | '+' '.join(f"{k}:{100*v/tot:.1f}" for k,v in byop.most_common(4)))
