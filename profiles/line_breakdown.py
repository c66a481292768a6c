"""Join an ncu SASS export (per-instruction executed counts / stall samples) with `nvdisasm -g` line info and
aggregate by source line.  usage: python profiles/line_breakdown.py <sass.csv> <disasm.txt> <mangled-substr> [ntop]"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, disasm, want = sys.argv[1], sys.argv[2], sys.argv[3]
ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1. line info per instruction index from nvdisasm
lines, cur, active = [], None, False
for ln in open(disasm):
    if ln.startswith('.text.'):
        active = want in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', ln):
        lines.append(cur)
# 2. per-instruction metrics from ncu (first kernel block)
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
num = lambda v: int(float(v)) if v not in ('', 'N/A') else 0
iex, ismp = hdr.index('Instructions Executed'), hdr.index('# Samples')
print('instructions: disasm %d, ncu %d' % (len(lines), len(data)))
n = min(len(lines), len(data))
ex, smp = defaultdict(int), defaultdict(int)
for i in range(n):
    ex[lines[i]] += num(data[i][iex])
    smp[lines[i]] += num(data[i][ismp])
te, ts = sum(ex.values()) or 1, sum(smp.values()) or 1
src = {}
def text(key):
    f, l = key
    if f not in src:
        import glob
        p = glob.glob('/root/repo/nnest_b200/csrc/**/' + f, recursive=True) + glob.glob('/usr/local/cuda/include/**/' + f, recursive=True)
        src[f] = open(p[0]).read().splitlines() if p else []
    return src[f][l - 1].strip()[:70] if 0 < l <= len(src[f]) else ''
print('total warp-instr %d, samples %d' % (te, ts))
for key in sorted(ex, key=lambda k: -(ex[k] / te + smp[k] / ts))[:ntop]:
    if key is None:
        continue
    print('%5.2f%% inst %5.2f%% smp  %-22s:%-4d %s' % (100.0 * ex[key] / te, 100.0 * smp[key] / ts, key[0], key[1], text(key)))
