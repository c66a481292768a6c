"""Summarise an `ncu --page source --print-source sass --csv` export: stall reasons, top instructions, opcode mix.
usage: python profiles/sass_summary.py <sass.csv> [kernel-substring] [ntop]"""
import csv
import sys
from collections import Counter

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ''
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(path)))
# split per kernel: a block starts with a "Kernel Name" row followed by a header row
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None and len(r) == len(cur['hdr']):
        cur['data'].append(r)
for b in blocks:
    if want not in b['name']:
        continue
    hdr, data = b['hdr'], b['data']
    isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
    num = lambda v: int(float(v)) if v not in ('', 'N/A') else 0
    tot = sum(num(r[isamp]) for r in data) or 1
    print('==', b['name'][:90], '| SASS instructions', len(data), '| samples', tot)
    agg = {hdr[i]: sum(num(r[i]) for r in data) for i in stall_cols}
    stot = sum(agg.values()) or 1
    print('-- stall reasons (share of stall samples)')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print('   %-26s %6.3f' % (k, v / stot))
    print('-- top instructions by samples')
    for r in sorted(data, key=lambda r: -num(r[isamp]))[:ntop]:
        rs = sorted([(num(r[i]), hdr[i][6:]) for i in stall_cols], reverse=True)[:2]
        print('   %5.2f%% ex=%9s %-62s %s' % (100.0 * num(r[isamp]) / tot, r[iex], r[isrc][:62], rs))
    c, s = Counter(), Counter()
    for r in data:
        toks = r[isrc].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        op = op.split('.')[0]
        c[op] += num(r[iex])
        s[op] += num(r[isamp])
    te = sum(c.values()) or 1
    print('-- opcode mix: executed share / sample share')
    for op, v in c.most_common(24):
        print('   %-12s %6.2f%% %6.2f%%' % (op, 100.0 * v / te, 100.0 * s[op] / tot))
    break
