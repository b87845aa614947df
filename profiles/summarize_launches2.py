"""Aggregate an ncu launch list (gpu__time_duration + dram bytes, --csv) by kernel.
usage: python profiles/summarize_launches2.py launches.csv [top]"""
import csv, collections, re, sys
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 30
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; data=rows[hi+1:]
ki,mi,ui,vi,gi=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Unit'),hdr.index('Metric Value'),hdr.index('Grid Size')
per=collections.OrderedDict()
for r in data:
    if len(r)<=vi: continue
    d=per.setdefault(r[0],{'k':r[ki],'g':r[gi]})
    v=float(r[vi].replace(',',''))
    u=r[ui]
    if u=='Kbyte': v*=1e3
    elif u=='Mbyte': v*=1e6
    elif u=='Gbyte': v*=1e9
    elif u in ('usecond','us'): v*=1e3
    elif u in ('msecond','ms'): v*=1e6
    elif u in ('second','s'): v*=1e9
    d[r[mi]]=v
agg=collections.OrderedDict(); tot=0
for id,d in per.items():
    n=d['k']
    key=re.sub(r'\(.*$','',n)
    key=key.replace('void ','').replace('tmk::','').replace('tc::','').replace('(int)','').replace('(bool)','')
    a=agg.setdefault(key,[0,0.,0.,0.])
    t=d.get('gpu__time_duration.sum',0)/1e3
    a[0]+=1; a[1]+=t; a[2]+=d.get('dram__bytes_read.sum',0); a[3]+=d.get('dram__bytes_write.sum',0)
    tot+=t
print(f"{'kernel':100s} {'n':>4s} {'total us':>10s} {'avg us':>9s} {'share':>6s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
for k,(c,t,rd,wr) in sorted(agg.items(), key=lambda x:-x[1][1])[:top]:
    print(f"{k[:100]:100s} {c:4d} {t:10.1f} {t/c:9.1f} {100*t/tot:5.1f}% {rd/1e6:11.1f} {wr/1e6:11.1f}")
print(f"total {tot:.1f} us over {len(per)} launches")
