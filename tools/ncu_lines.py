"""Stall samples per CUDA source line for one kernel of an .ncu-rep.
Maps ncu's per-SASS-instruction samples to source lines through `nvdisasm -g` of the cubin
extracted from the built library (instruction order is identical).
Usage: python tools/ncu_lines.py report.ncu-rep <kernel-substring> [topn]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, ksub = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'prosper_b200', 'lib', 'libprosper_b200.so')
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern = None
cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        if ksub in r[1] and kern is None:
            kern = cur
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and r:
        cur['rows'].append(r)
h = kern['hdr']; si = h.index('# Samples'); ii = h.index('Instructions Executed')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, capture_output=True)
# find the mangled function by matching instruction count
best = None
for f in os.listdir(tmp):
    if not f.endswith('.cubin') or '-' in f:
        continue
    txt = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    funcs = re.split(r'\n\s*\.section\s+\.text\.', txt)
    for fn in funcs[1:]:
        name = fn.split(',')[0].split()[0]
        lines = fn.split('\n')
        ins = []
        curline = None
        curfile = None
        for ln in lines:
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                curfile, curline = os.path.basename(m.group(1)), int(m.group(2)); continue
            if re.match(r'\s*/\*[0-9a-f]{4,}\*/', ln):
                ins.append((curfile, curline))
        # trailing padding instructions may be missing from ncu's listing: take the closest function within 16
        diff = abs(len(ins) - len(kern['rows']))
        if diff <= 16 and ksub.split('<')[0] in name and (best is None or diff < best[2]):
            best = (name, ins, diff)
if best is not None:
    best = best[:2]
if best is None:
    print('could not match function by instruction count', len(kern['rows'])); sys.exit(1)
name, ins = best
agg = {}; execs = {}
tot = 0
for (f, l), r in zip(ins, kern['rows']):
    v = float(r[si]); tot += v
    agg[(f, l)] = agg.get((f, l), 0) + v
    execs[(f, l)] = execs.get((f, l), 0) + float(r[ii])
print(kern['name'], 'samples', tot)
src_cache = {}
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1])[:topn]:
    text = ''
    p = os.path.join(ROOT, 'prosper_b200', 'csrc', f or '')
    if f and os.path.exists(p):
        src_cache.setdefault(p, open(p).read().split('\n'))
        if l and l <= len(src_cache[p]):
            text = src_cache[p][l - 1].strip()[:100]
    print('%5.1f%%  %-14s:%-4s inst %-10d %s' % (100 * v / tot, f, l, execs[(f, l)], text))

if len(sys.argv) > 4:
    # group by line ranges: "name:lo-hi,name:lo-hi"
    groups = []
    for g in sys.argv[4].split(','):
        nm, rg = g.split(':'); lo, hi = rg.split('-'); groups.append((nm, int(lo), int(hi)))
    gs = {}; gi = {}
    ti = sum(execs.values())
    for (f, l), v in agg.items():
        key = 'other(' + str(f) + ')'
        if l:
            for nm, lo, hi in groups:
                if lo <= l <= hi:
                    key = nm
        gs[key] = gs.get(key, 0) + v; gi[key] = gi.get(key, 0) + execs[(f, l)]
    print('total warp instructions', ti)
    for k, v in sorted(gs.items(), key=lambda x: -x[1]):
        print('%5.1f%% samples  %5.1f%% instr  %s' % (100 * v / tot, 100 * gi[k] / ti, k))
