import csv, collections, re, sys
path = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows=list(csv.reader(open(path)))
hdr=None; data=[]
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr=r
        continue
    data.append(dict(zip(hdr,r)))
half=data[int(len(data)*frac):]
def ms(d):
    v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
    return v/1e6 if u in ('ns','nsecond') else v/1e3 if u in('us','usecond') else v
agg=collections.defaultdict(lambda:[0,0.0]); tot=0
for d in half:
    name=re.sub(r'\(.*','',d['Kernel Name']); name=re.sub(r'^void ','',name)
    agg[name][0]+=1; agg[name][1]+=ms(d); tot+=ms(d)
print('launches',len(half),'sum ms %.3f'%tot)
for k,(c,m) in sorted(agg.items(), key=lambda x:-x[1][1])[:int(sys.argv[3]) if len(sys.argv)>3 else 25]:
    print(f"{m:8.3f} ms {100*m/tot:5.1f}% {c:4d}  {k[:90]}")
