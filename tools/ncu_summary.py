"""Summarise an .ncu-rep: key raw metrics + the hottest source lines (needs -lineinfo).
Usage: python tools/ncu_summary.py report.ncu-rep [n_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_op_dmma.sum', 'sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.avg', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    print("== kernel:", r[hdr.index('Kernel Name')][:100] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, rows[1], r):
        if h in KEYS or (h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio')) \
                or (h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')):
            try:
                if float(v.replace(',', '')) == 0:
                    continue
            except ValueError:
                pass
            print("  %-90s %s %s" % (h, v, u))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if rows:
    # find header row
    for i, r in enumerate(rows):
        if 'Source' in r and any('Sampling' in c for c in r):
            hdr = r
            body = rows[i + 1:]
            break
    else:
        hdr, body = None, []
    if hdr:
        si = hdr.index('Source')
        ci = [j for j, c in enumerate(hdr) if c.startswith('# Samples') or c == 'Warp Stall Sampling (All Samples)' or c == 'Warp Stall Sampling (All Cycles)']
        ci = ci[0] if ci else None
        ii = hdr.index('Instructions Executed') if 'Instructions Executed' in hdr else None
        li = hdr.index('#') if '#' in hdr else 0
        if ci is not None:
            tot = 0
            items = []
            for r in body:
                try:
                    v = float(r[ci])
                except (ValueError, IndexError):
                    continue
                tot += v
                items.append((v, r[li], r[si].strip()[:110], r[ii] if ii is not None else ''))
            items.sort(reverse=True)
            print("== hottest source lines (stall samples, %% of %d)" % tot)
            for v, ln, s, ie in items[:topn]:
                print("  %5.1f%%  line %-5s inst %-10s %s" % (100 * v / max(tot, 1), ln, ie, s))
