"""Per barrier-separated phase of an `ncu --page source --csv` export: stall samples by opcode class, and the
instructions that collect the most samples.  usage: python tools/ncu_phase_ops.py <source.csv> [phase ...]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
want = [int(a) for a in sys.argv[2:]]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def num(s):
    try: return float(s)
    except Exception: return 0.0
def opclass(src):
    t = src.split()
    op = t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '')
    for k in ('DFMA', 'DMUL', 'DADD', 'LDS', 'STS', 'LDGSTS', 'LDG', 'STG', 'LDL', 'STL', 'BAR', 'FSEL', 'MUFU', 'LDC', 'SHFL'):
        if op.startswith(k): return 'fp64' if k in ('DFMA', 'DMUL', 'DADD') else k
    return 'other'
phase = 0
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: collections.defaultdict(float)))
top = collections.defaultdict(list)
for i, r in enumerate(data):
    src = r[ix['Source']].strip()
    if 'BAR.SYNC' in src: phase += 1
    c = opclass(src)
    n = num(r[ix['# Samples']])
    agg[phase][c]['samples'] += n
    agg[phase][c]['inst'] += num(r[ix['Instructions Executed']])
    for s in stalls: agg[phase][c][s] += num(r[ix[s]])
    top[phase].append((n, i, src))
for p in sorted(agg):
    if want and p not in want: continue
    tot = sum(v['samples'] for v in agg[p].values())
    print('== phase %d  samples %d' % (p, tot))
    for c, v in sorted(agg[p].items(), key=lambda kv: -kv[1]['samples']):
        if v['samples'] < 0.01 * tot: continue
        ts = sorted([(v[s], s[6:]) for s in stalls], reverse=True)[:4]
        print('   %-7s %5.1f%%  inst %7.2fM  %s' % (c, 100 * v['samples'] / tot, v['inst'] / 1e6,
                                                  ' '.join('%s=%.0f%%' % (s, 100 * x / max(v['samples'], 1)) for x, s in ts)))
    if want:
        for n, i, src in sorted(top[p], reverse=True)[:25]:
            r = data[i]
            ts = sorted([(num(r[ix[s]]), s[6:]) for s in stalls], reverse=True)[:2]
            print('      %5d  #%-5d %-58s %s' % (n, i, src[:58], ' '.join('%s=%d' % (s, x) for x, s in ts if x)))
