"""Joins an `ncu --page source --csv` export with `nvdisasm -g -c` line info of the same build: share of
executed instructions / stall samples / threads per instruction by source line.
    python tools/ncu_linemap.py <sass from nvdisasm> <.text.label of the kernel> <source csv> [top n]"""
import csv,re,sys
from collections import defaultdict
sass_file, kern_label, src_csv = sys.argv[1], sys.argv[2], sys.argv[3]
topn=int(sys.argv[4]) if len(sys.argv)>4 else 45
lines=open(sass_file).read().split('\n')
start=next(i for i,l in enumerate(lines) if l.startswith(kern_label+':'))
addr2line={}
cur=None
for l in lines[start+1:]:
    if l.startswith('.text.') or l.startswith('//-----'): 
        if addr2line: break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);',l)
    if m: addr2line[int(m.group(1),16)]=(cur,m.group(2))
rows=list(csv.reader(open(src_csv)))
h=rows[1]; data=rows[2:]
ia=h.index('Address'); ins=h.index('Instructions Executed'); ismp=h.index('# Samples'); ith=h.index('Thread Instructions Executed')
base=int(data[0][ia],16) if data[0][ia].startswith('0x') or re.match('^[0-9a-f]+$',data[0][ia]) else 0
agg=defaultdict(lambda:[0,0,0])
tot=tots=0
for k,r in enumerate(data):
    a=k*16
    ln=addr2line.get(a,(None,''))[0]
    e=int(r[ins]); s=int(r[ismp]); t=int(r[ith])
    agg[ln][0]+=e; agg[ln][1]+=s; agg[ln][2]+=t
    tot+=e; tots+=s
print('total inst',tot,'samples',tots,'mapped addrs',len(addr2line),'rows',len(data))
for ln,(e,s,t) in sorted(agg.items(),key=lambda kv:-kv[1][0])[:topn]:
    print(f"{str(ln):38s} inst {e/tot*100:5.2f}%  samples {s/tots*100:5.2f}%  thr/inst {t/max(e,1):4.1f}")
# by file
byf=defaultdict(lambda:[0,0])
for ln,(e,s,t) in agg.items():
    f=ln[0] if ln else None
    byf[f][0]+=e; byf[f][1]+=s
print({f:(round(v[0]/tot*100,1),round(v[1]/tots*100,1)) for f,v in byf.items()})
