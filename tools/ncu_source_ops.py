"""Aggregate an `ncu --page source --csv` dump per SASS opcode: executed warp-instructions,
shared-memory wavefronts (and their conflict-free ideal) and stall samples, all per 32 keys.
usage: ncu -i x.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_ops.py src.csv <keys per launch>"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
keys = float(eval(sys.argv[2])) if len(sys.argv) > 2 else 2.0**28
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {'name': r[1], 'rows': []}; secs.append(cur); continue
    if r and r[0] == "Address": cur['hdr'] = r; continue
    if cur is not None and len(r) > 5: cur['rows'].append(r)
for s in secs:
    h = s['hdr']; ix = {n: i for i, n in enumerate(h)}
    print(s['name'][:140], len(s['rows']), "SASS instructions")
    tot = collections.defaultdict(lambda: [0, 0, 0, 0])
    N = keys / 32
    for r in s['rows']:
        op = [o for o in r[ix['Source']].split() if not o.startswith('@')][0]
        if not op.startswith(('LDS', 'STS', 'ATOMS', 'LDG', 'STG', 'RED')): op = op.split('.')[0]
        t = tot[op]
        t[0] += int(r[ix['Instructions Executed']]); t[1] += int(r[ix['L1 Wavefronts Shared']])
        t[2] += int(r[ix['L1 Wavefronts Shared Ideal']]); t[3] += int(r[ix['# Samples']])
    ti = sum(t[0] for t in tot.values()); ts = sum(t[3] for t in tot.values()); tw = sum(t[1] for t in tot.values())
    for op, t in sorted(tot.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"  {op:28s} inst/32keys {t[0]/N:7.2f}  shared wavefronts/32keys {t[1]/N:6.2f} (ideal {t[2]/N:5.2f})  stall samples {100*t[3]/max(ts,1):5.1f}%")
    print(f"  TOTAL inst/32keys {ti/N:.2f}  shared wavefronts/32keys {tw/N:.2f}")
