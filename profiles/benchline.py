import sys, json
tag = sys.argv[1]
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(tag, '%.3f G/s' % (d['value'] / 1e9), '%.3f ms' % d['ms_per_step'], 'e2e %.3f' % (d['e2e']['value'] / 1e9), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'frac %.3f' % d['roofline']['frac'], 'kernel %.3f ms' % d['roofline']['launch_ms'])
