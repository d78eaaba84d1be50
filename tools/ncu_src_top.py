"""Summarise an `ncu --page source --csv` export: instructions with most stall samples."""
import csv, sys
path=sys.argv[1]; n=int(sys.argv[2]) if len(sys.argv)>2 else 30
rows=list(csv.reader(open(path)))
his=[i for i,r in enumerate(rows) if r and r[0]=='Address']
for bi,hi in enumerate(his):
    hdr=rows[hi]; idx={h:i for i,h in enumerate(hdr)}
    end=his[bi+1]-1 if bi+1<len(his) else len(rows)
    data=[r for r in rows[hi+1:end] if len(r)==len(hdr) and r[0]!='Address']
    def I(r,k):
        try: return int(float(r[idx[k]] or 0))
        except: return 0
    tot=sum(I(r,'# Samples') for r in data)
    print('=== block',bi,rows[hi-1][:2] if hi>0 else '', 'total samples',tot,'n instr',len(data))
    keys=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg={k:sum(I(r,k) for r in data) for k in keys}
    print({k:v for k,v in sorted(agg.items(),key=lambda x:-x[1]) if v})
    top=sorted(data,key=lambda r:-I(r,'# Samples'))[:n]
    for r in sorted(top,key=lambda r:r[idx['Address']]):
        st={k[6:]:I(r,k) for k in keys if I(r,k)}
        st=dict(sorted(st.items(),key=lambda x:-x[1])[:3])
        print(r[idx['Address']][-5:], str(I(r,'# Samples')).rjust(6), str(I(r,'Instructions Executed')).rjust(9), r[idx['Source']][:64].ljust(64), st)
