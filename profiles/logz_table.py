"""Tabulate the matched-settings evidence runs (profiles/r2_data/{ref,repo}_logz.jsonl): per setting and implementation the
number of seeds, mean logZ, standard deviation over seeds, standard error of the mean and the nominal sqrt(H/nlive).
usage: python profiles/logz_table.py"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rows = []
for fn in ('ref_logz.jsonl', 'repo_logz.jsonl'):
    rows += [json.loads(l) for l in open(os.path.join(HERE, 'r2_data', fn)) if l.strip()]
tags = []
for r in rows:
    if r['tag'] not in tags and r['tag'] != 'smoke':
        tags.append(r['tag'])
print('| setting | impl | seeds | mean logZ | sd over seeds | s.e.m. | nominal sqrt(H/nlive) | niter | ncall | wall s / run |')
print('|---|---|---|---|---|---|---|---|---|---|')
for tag in tags:
    for impl in ('reference', 'nnest_b200'):
        a = [r for r in rows if r['tag'] == tag and r['impl'] == impl]
        if not a:
            continue
        z = np.array([r['logz'] for r in a])
        sd = z.std(ddof=1) if len(z) > 1 else float('nan')
        print('| %s | %s | %d | %.3f | %.3f | %.3f | %.3f | %.0f | %.3g | %.1f |' % (
            tag, impl, len(a), z.mean(), sd, sd / np.sqrt(len(z)), np.mean([r['logzerr'] for r in a]),
            np.mean([r['niter'] for r in a]), np.mean([r['ncall'] for r in a]), np.mean([r['wall_s'] for r in a])))
