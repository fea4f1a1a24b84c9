"""Aggregate an `ncu --page source --csv` export per barrier-separated phase of a kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def num(s):
    try: return float(s)
    except Exception: return 0.0
phase = 0; agg = {}; tot = 0
for r in data:
    src = r[ix['Source']]
    if 'BAR.SYNC' in src: phase += 1
    a = agg.setdefault(phase, {'samples': 0, 'inst': 0, 'fp64': 0, 'lds': 0, 'sts': 0, 'ldg': 0, 'stg': 0, 'ldl': 0, **{s: 0 for s in stalls}})
    n = num(r[ix['# Samples']]); a['samples'] += n; tot += n
    ie = num(r[ix['Instructions Executed']]); a['inst'] += ie
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '')
    for key, pre in (('fp64', ('DFMA', 'DADD', 'DMUL')), ('lds', ('LDS',)), ('sts', ('STS',)), ('ldg', ('LDG',)), ('stg', ('STG',)), ('ldl', ('LDL', 'STL'))):
        if op.startswith(pre): a[key] += ie
    for s in stalls: a[s] += num(r[ix[s]])
print('total samples', tot)
for p, a in agg.items():
    top = sorted([(a[s], s) for s in stalls], reverse=True)[:4]
    print(p, 'samples %5.1f%%' % (100 * a['samples'] / tot), 'inst %.1fM fp64 %.1fM lds %.2fM sts %.2fM ldg %.2fM stg %.2fM ldl %.2fM' % tuple(a[k] / 1e6 for k in ('inst', 'fp64', 'lds', 'sts', 'ldg', 'stg', 'ldl')), ' '.join('%s=%.0f%%' % (s[6:], 100 * v / max(a['samples'], 1)) for v, s in top))
