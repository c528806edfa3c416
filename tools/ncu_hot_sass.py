"""Top SASS instructions by stall samples from an .ncu-rep, with dominant stall reason.
Usage: python tools/ncu_hot_sass.py report.ncu-rep [topn] [kernel-index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# split per kernel
kernels = []
cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}; kernels.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and r:
        cur['rows'].append(r)
ki = int(sys.argv[3]) if len(sys.argv) > 3 else 0
k = kernels[ki]
h = k['hdr']
si = h.index('# Samples'); ii = h.index('Instructions Executed'); src = h.index('Source')
stall_cols = [j for j, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(float(r[si]) for r in k['rows'])
print(k['name'], 'total samples', tot, 'instructions', len(k['rows']))
agg = {}
for j in stall_cols:
    agg[h[j]] = sum(float(r[j]) for r in k['rows'])
print('stall totals:', {a: round(100 * b / tot, 1) for a, b in sorted(agg.items(), key=lambda x: -x[1])[:8]})
idx = sorted(range(len(k['rows'])), key=lambda i: -float(k['rows'][i][si]))[:topn]
for i in sorted(idx):
    r = k['rows'][i]
    st = max(stall_cols, key=lambda j: float(r[j]))
    print("%5d %5.1f%% exec %-9s %-16s %s" % (i, 100 * float(r[si]) / tot, r[ii], h[st], r[src].strip()[:90]))
