import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ia=hdr.index('Address'); isrc=hdr.index('Source'); isamp=hdr.index('# Samples'); ie=hdr.index('Instructions Executed')
stall_cols=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
recs=[]
for r in data:
    if not r[ia].startswith('0x'): continue
    st={h:int(r[i] or 0) for i,h in stall_cols if int(r[i] or 0)>0}
    recs.append((int(r[ia],16), r[isrc], int(r[isamp] or 0), int(r[ie] or 0), st))
base=recs[0][0]
def region(lo,hi,name):
    tot=collections.Counter(); n=0; ins=0
    for a,src,s,e,st in recs:
        if lo<=a-base<hi:
            n+=s; ins+=e
            for h,v in st.items(): tot[h]+=v
    print(name, 'samples',n,'inst',ins, ' '.join(f'{h[6:]}:{v}' for h,v in tot.most_common(8)))
def listing(lo,hi,thr):
    for a,src,s,e,st in recs:
        if lo<=a-base<hi and s>=thr:
            print(hex(a-base), s, 'x',e, src[:60], dict(sorted(st.items(), key=lambda kv:-kv[1])[:2]))
if __name__=='__main__':
    for spec in sys.argv[2:]:
        lo,hi,name=spec.split(':'); region(int(lo,16),int(hi,16),name)
