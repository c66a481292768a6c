"""Opcode counts per kernel of libnnb.so: python profiles/sass_opcodes.py > profiles/r2_sass_summary.txt
(cuobjdump -sass nnest_b200/lib/libnnb.so; the mnemonics of the B200_PROFILING.md table)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'nnest_b200', 'lib', 'libnnb.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
names = {}
try:
    dem = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', out)), capture_output=True, text=True).stdout.split('\n')
    names = dict(zip(re.findall(r'Function : (\S+)', out), dem))
except Exception:
    pass
GROUPS = [('UTCHMMA', r'UTC\w*MMA'), ('LDTM', r'LDTM'), ('STTM', r'STTM'), ('UTCBAR', r'UTCBAR'), ('UBLKCP', r'UBLKCP'),
          ('SYNCS', r'SYNCS'), ('FFMA', r'FFMA'), ('DFMA', r'DFMA'), ('MUFU', r'MUFU'), ('BAR', r'BAR\b'), ('RED', r'RED\b'),
          ('ATOM', r'ATOM'), ('LDS', r'LDS'), ('STS', r'STS'), ('LDC', r'LDCU?\b'), ('LDG', r'LDG'), ('STG', r'STG'), ('SHFL', r'SHFL')]
print('# SASS opcode evidence of nnest_b200/lib/libnnb.so (sm_100a), round 2.  Made by profiles/sass_opcodes.py: cuobjdump -sass')
print('# libnnb.so, counting the mnemonics of the B200_PROFILING.md table per kernel: tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->')
print('# LDTM/STTM, tcgen05.commit -> UTCBAR, TMA bulk copy (cp.async.bulk) -> UBLKCP, mbarrier -> SYNCS.*, constant-bank loads of')
print('# the kernel-parameter constants -> LDC/LDCU; FFMA / MUFU for the FP32 kernels.')
print('%-110s %6s  %s' % ('kernel', 'SASS', 'counts'))
rows = []
for m in re.finditer(r'Function : (\S+)\n(.*?)(?=\n\s*Function : |\Z)', out, re.S):
    fn, body = m.group(1), m.group(2)
    ops = re.findall(r'^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', body, re.M)
    if not ops:
        continue
    c = collections.Counter()
    for o in ops:
        for g, pat in GROUPS:
            if re.match(pat, o):
                c[g] += 1
                break
    rows.append((len(ops), names.get(fn, fn)[:110], ' '.join('%s=%d' % (g, c[g]) for g, _ in GROUPS if c[g])))
for n, name, cs in sorted(rows, reverse=True):
    print('%-110s %6d  %s' % (name, n, cs))
