"""Aggregate ncu --page source --csv stall samples per SASS instruction / opcode (dev aid)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        blocks.append(cur)
    elif r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and r:
        cur['rows'].append(r)
for b in blocks:
    h = b['hdr']
    ia, isamp = h.index('Source'), h.index('Warp Stall Sampling (All Samples)')
    tot = sum(int(r[isamp] or 0) for r in b['rows'])
    print(b['name'], 'instr', len(b['rows']), 'samples', tot)
    for k, r in enumerate(b['rows']):
        r.append(k)
    for r in sorted(b['rows'], key=lambda r: -int(r[isamp] or 0))[:topn]:
        k = r[-1]
        prev = b['rows'][k - 1][ia].strip()[:50] if k > 0 else ''
        print('   %6d %5.1f%%  #%5d %-60s | prev: %s' % (int(r[isamp]), 100 * int(r[isamp]) / tot, k, r[ia].strip()[:60], prev))
    agg = collections.Counter()
    for r in b['rows']:
        t = r[ia].strip().split()
        op = t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '')
        agg[op.split('.')[0]] += int(r[isamp] or 0)
    print('   by opcode:', [(k, round(100 * v / tot, 1)) for k, v in agg.most_common(14)])
