import csv, sys, collections, re, bisect
path=sys.argv[1]
rows=csv.reader(open(path))
fname=None; hdr=None
per_line=collections.Counter(); per_line_inst=collections.Counter(); stall=collections.defaultdict(collections.Counter)
src={}
for r in rows:
    if len(r)==2 and r[0]=="File Path": fname=r[1].split('/')[-1]; continue
    if len(r)>10 and r[0]=="Line No": hdr=r; idx={h:i for i,h in enumerate(hdr)}; si=hdr.index("# Samples"); ii=hdr.index("Instructions Executed"); continue
    if hdr and len(r)==len(hdr) and r[2]=="-":   # source line aggregate row
        try: s=int(r[si]); ie=int(r[ii])
        except: continue
        key=(fname,int(r[0])); per_line[key]+=s; per_line_inst[key]+=ie; src[key]=r[1].strip()[:100]
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                try: stall[key][h]+=int(r[idx[h]])
                except: pass
tot=sum(per_line.values()); toti=sum(per_line_inst.values())
print("total samples",tot,"total inst",toti)
# per function
def funcs(p):
    out=[]
    try: L=open(p).read().split('\n')
    except: return out
    for i,l in enumerate(L,1):
        if l and not l[0].isspace() and not l.startswith('//') and not l.startswith('#') and not l.startswith('}'):
            m=re.search(r'\b([a-zA-Z_0-9]+)\(', l)
            if m and (l.rstrip().endswith('{') or l.rstrip().endswith(',')): out.append((i,m.group(1)))
    return out
ftot=collections.Counter(); fi=collections.Counter(); fst=collections.defaultdict(collections.Counter)
cache={}
for (f,ln),s in per_line.items():
    if f not in cache: cache[f]=funcs('/root/repo/moby_b200/csrc/'+f)
    fl=cache[f]; starts=[a for a,_ in fl]; k=bisect.bisect_right(starts,ln)-1
    name=(f, fl[k][1] if k>=0 else '?')
    ftot[name]+=s; fi[name]+=per_line_inst[(f,ln)]
    for h,c in stall[(f,ln)].items(): fst[name][h]+=c
print("--- by function")
for name,s in ftot.most_common(28):
    top=", ".join(f"{h[6:]}={c}" for h,c in fst[name].most_common(3))
    print(f"{100*s/tot:5.1f}% samples {100*fi[name]/toti:5.1f}% inst  {name[0]}:{name[1]}   [{top}]")
print("--- by line")
for key,s in per_line.most_common(25):
    top=", ".join(f"{h[6:]}={c}" for h,c in stall[key].most_common(2))
    print(f"{100*s/tot:5.1f}% {per_line_inst[key]:9d} {key[0]}:{key[1]} {src[key]}  [{top}]")
