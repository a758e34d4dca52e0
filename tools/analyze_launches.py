import csv, collections, re, sys, json
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
lines=[l for l in open(path) if not l.startswith('==')]
rows=[]
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum':
        v=float(row['Metric Value'].replace(',',''))
        if row['Metric Unit'] in('nsecond','ns'): v/=1e3
        rows.append((int(row['ID']),re.sub(r'\(.*','',row['Kernel Name']).replace('void ','').replace('fb200::',''),v,row.get('Grid Size')))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 235
step=rows[:n]
tot=sum(r[2] for r in step)
print(len(rows),'launches captured; step of',len(step),'total us',round(tot,1))
agg=collections.defaultdict(lambda:[0,0.0])
for _,nm,v,g in step: agg[nm[:75]][0]+=1; agg[nm[:75]][1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:16]:
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={c:3d}  {k}")
print('top single launches:')
for r in sorted(step,key=lambda r:-r[2])[:12]:
    print(f"{r[2]:9.1f} us id={r[0]} grid={r[3]} {r[1][:70]}")
