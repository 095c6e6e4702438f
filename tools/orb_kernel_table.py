#!/usr/bin/env python
"""Per-kernel table (time, share, warp instructions, issue-slot and warp occupancy) from an ncu --csv log of the ORB front end.
usage: python tools/orb_kernel_table.py <orb_metrics.csv>"""
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[start+1:]:
    if len(r)<=vi: continue
    try: agg[r[ki].split('(')[0].replace('<unnamed>::','')][r[mi]].append(float(r[vi].replace(',','')))
    except: pass
# warmup run included: two calls -> halve
tot=sum(sum(m.get('gpu__time_duration.sum',[0])) for m in agg.values())
print("total ms (all captured launches)",tot/1e6)
for k,m in sorted(agg.items(), key=lambda kv:-sum(kv[1].get('gpu__time_duration.sum',[0]))):
    t=sum(m.get('gpu__time_duration.sum',[0])); n=len(m.get('gpu__time_duration.sum',[]))
    print(f"{k[:34]:34s} n={n:3d} t={t/1e6:7.3f} ms {100*t/tot:5.1f}%  inst={sum(m.get('smsp__inst_executed.sum',[0]))/1e6:9.1f}M issue={sum(m.get('smsp__issue_active.avg.pct_of_peak_sustained_active',[0]))/max(n,1):5.1f} warps={sum(m.get('sm__warps_active.avg.pct_of_peak_sustained_active',[0]))/max(n,1):5.1f}")
