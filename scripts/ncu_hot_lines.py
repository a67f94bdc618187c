import re,csv,sys,collections
kern=sys.argv[1]; sasscsv=sys.argv[2]
lines=open('/tmp/dis.txt').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith(kern+':')][0]
addr2line={}; cur=None; curfile=None
for l in lines[start:]:
    if l.startswith('//-----') and 'text.' in l and kern not in l: break
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        curfile=m.group(1).split('/')[-1]; cur=int(m.group(2)); inl=m.group(3); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m and cur is not None:
        addr2line[int(m.group(1),16)]=(curfile,cur)
rows=list(csv.reader(open(sasscsv)))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
agg=collections.defaultdict(lambda:[0,0,0])
base=None
for r in rows[2:]:
    a=int(r[ix['Address']],16)
    if base is None: base=a
    key=addr2line.get(a-base,('?',0))
    w=float(r[ix['Instructions Executed']] or 0); t=float(r[ix['Thread Instructions Executed']] or 0); s=float(r[ix['# Samples']] or 0)
    agg[key][0]+=w; agg[key][1]+=t; agg[key][2]+=s
tw=sum(v[0] for v in agg.values()); ts=sum(v[2] for v in agg.values())
src={}
for f in ('qk_kernels.cuh','qk_device.cuh'):
    src[f]=open('/root/repo/quokkasim_b200/csrc/'+f).read().split('\n')
print("total warp inst %.3e samples %d"%(tw,ts))
for k,v in sorted(agg.items(), key=lambda x:-x[1][2])[:int(sys.argv[3])]:
    f,l=k
    text=src.get(f,[''])[l-1].strip()[:90] if f in src and l>0 else ''
    print("%-16s %4d  inst %5.2f%%  samp %5.2f%%  lanes %4.1f  | %s"%(f,l,100*v[0]/tw,100*v[2]/ts,(v[1]/v[0] if v[0] else 0),text))
