import csv, re, sys
fn = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 20
with open(fn) as f:
    lines=[l for l in f if not l.startswith('==')]
rows=list(csv.DictReader(lines))
def short(n):
    n=re.sub(r'\(.*','',n); return n.replace('psd::','').replace('void ','')[:60]
idx=[i for i,r in enumerate(rows) if 'transform_kernel' in r['Kernel Name']]
start,end=idx[-2],idx[-1]
tot=0
for r in rows[start:end]:
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
    tot+=v
    if v>thr: print(f"{v:9.1f} us  {short(r['Kernel Name'])} grid={r['Grid Size']} blk={r['Block Size']}")
print("step total", tot, "launches", end-start)
