"""Static SASS opcode histogram of one kernel: python tools/sass_hist.py <object-or-cubin> <kernel-substring> [per-line]"""
import collections, re, subprocess, sys
obj, sub = sys.argv[1], sys.argv[2]
txt = subprocess.run(['nvdisasm', '-g', obj], capture_output=True, text=True).stdout if obj.endswith('.cubin') else None
if txt is None:
    import tempfile, os
    d = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, capture_output=True)
    txt = ''.join(subprocess.run(['nvdisasm', '-g', os.path.join(d, f)], capture_output=True, text=True).stdout for f in os.listdir(d) if f.endswith('.cubin'))
for fn in re.split(r'\n\s*\.section\s+\.text\.', txt)[1:]:
    name = fn.split(',')[0].split()[0]
    if sub not in name:
        continue
    cur = None; ops = collections.Counter(); byline = collections.Counter(); n = 0
    for ln in fn.split('\n'):
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            t = re.sub(r'^@!?U?P\d+\s+', '', m.group(2)); ops[t.split()[0].split('.')[0]] += 1; byline[cur] += 1; n += 1
    print(name[:70], n)
    print(' ', ops.most_common(22))
    if len(sys.argv) > 3:
        print(' ', byline.most_common(14))
