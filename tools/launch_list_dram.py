"""ncu launch list with DRAM bytes (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
of tools/profile_step.py N 2 -> markdown table of the SECOND EM iteration (launches from the second transpose_w_kernel
on) and the traffic record bench.py reads.
Usage: python tools/launch_list_dram.py launches.csv out.md out_traffic.json"""
import csv, json, re, sys
path, out_md, out_json = sys.argv[1:4]
lines = [l for l in open(path) if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
idi, ki, mi, ui, vi = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
launches = {}
order = []
for r in rd:
    if len(r) <= vi:
        continue
    i = int(r[idi])
    if i not in launches:
        launches[i] = {'name': re.sub(r"\(.*", "", r[ki]).replace("pet::", "")}
        order.append(i)
    v = float(r[vi].replace(",", ""))
    u = r[ui].lower()
    if r[mi] == 'gpu__time_duration.sum':
        launches[i]['ms'] = v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else v)
    else:
        mult = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[u]
        launches[i]['rd' if 'read' in r[mi] else 'wr'] = v * mult
starts = [i for i in order if launches[i]['name'].endswith('transpose_w_kernel')]
first = starts[1] if len(starts) > 1 else starts[0]
sel = [launches[i] for i in order if i >= first]
agg = {}
for l in sel:
    a = agg.setdefault(l['name'], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += l.get('ms', 0); a[2] += l.get('rd', 0); a[3] += l.get('wr', 0)
tot = sum(a[1] for a in agg.values())
rd_t, wr_t = sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())
with open(out_md, 'w') as f:
    f.write("| kernel | launches | total ms | share | DRAM read GB | DRAM write GB |\n|---|---:|---:|---:|---:|---:|\n")
    for k, (n, ms, r_, w_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.1f%% | %.2f | %.2f |\n" % (k, n, ms, 100 * ms / tot, r_ / 1e9, w_ / 1e9))
    f.write("\nTotal %.1f ms over %d launches; DRAM traffic of the step %.1f GB read + %.1f GB written = %.1f GB.\n"
            % (tot, len(sel), rd_t / 1e9, wr_t / 1e9, (rd_t + wr_t) / 1e9))
g = [l for l in sel if 'gemm_kernel' in l['name'] and 'oz::' in l['name']]
json.dump({"oz_gemm_bytes_per_launch": sum(l.get('rd', 0) + l.get('wr', 0) for l in g) / max(1, len(g)),
           "oz_gemm_launches": len(g), "step_dram_bytes": rd_t + wr_t, "step_dram_read": rd_t, "step_dram_write": wr_t,
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                     "python tools/profile_step.py 1000000 2 (second iteration)"}, open(out_json, 'w'), indent=1)
print(open(out_md).read())
