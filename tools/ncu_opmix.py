"""Dynamic opcode mix of one kernel of an .ncu-rep (warp instructions executed per opcode).
Usage: python tools/ncu_opmix.py report.ncu-rep <kernel-substring> <units> [per-line]   (units: divide counts by this)"""
import collections, csv, io, re, subprocess, sys
rep, ksub, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern = None; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        if ksub in r[1] and kern is None:
            kern = cur
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and r:
        cur['rows'].append(r)
h = kern['hdr']; ii = h.index('Instructions Executed'); so = h.index('Source')
ops = collections.Counter(); tot = 0
for r in kern['rows']:
    n = int(r[ii]); t = re.sub(r'^@!?U?P\d+\s+', '', r[so]); ops[t.split()[0].split('.')[0]] += n; tot += n
print(kern['name'][:80], "total warp instructions", tot, "=", round(tot / units, 1), "per unit")
for k, v in ops.most_common(24):
    print("  %-10s %9.1f  %5.1f%%" % (k, v / units, 100 * v / tot))
