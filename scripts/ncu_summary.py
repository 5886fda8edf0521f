"""Summarise an .ncu-rep: per kernel key metrics + stall mix + opcode mix.  usage: ncu_summary.py rep [rows]"""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]; nrows = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0]
    if name in seen:
        continue
    seen.add(name)
    print('===', name)
    for w in want:
        if w in idx:
            print(f"  {w:62s} {r[idx[w]]:>16s} {units[idx[w]]}")
    if nrows:
        print(f"  thread-instr/row {float(r[idx['smsp__inst_executed.sum']]) * 32 / nrows:.1f}   "
              f"GB/s(alg) n/a   dram B/row {(float(r[idx['dram__bytes_read.sum']])) * (1e9 if units[idx['dram__bytes_read.sum']]=='Gbyte' else 1e6) / nrows:.1f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                          "regex:" + name.replace("void ", "").split('<')[0],
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    if len(srows) < 3:
        continue
    sh = srows[1]; si = {h: i for i, h in enumerate(sh)}
    data, seenaddr = [], set()
    for x in srows[2:]:
        if len(x) >= len(sh) - 2 and x[0].startswith('0x') and x[0] not in seenaddr:
            seenaddr.add(x[0]); data.append(x)
    tot = sum(int(x[si['# Samples']]) for x in data) or 1
    stalls = [h for h in sh if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {s: sum(int(x[si[s]] or 0) for x in data) for s in stalls}
    print("  SASS lines", len(data), " stalls:", ", ".join(f"{s[6:]} {100*v/tot:.0f}%" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    ex = sum(int(x[si['Instructions Executed']]) for x in data) or 1
    c = Counter()
    for x in data:
        parts = x[si['Source']].split()
        op = parts[1] if parts[0].startswith('@') else parts[0]
        c[op.split('.')[0]] += int(x[si['Instructions Executed']])
    print("  opcodes:", ", ".join(f"{k} {100*v/ex:.0f}%" for k, v in c.most_common(12)))
