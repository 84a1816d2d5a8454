"""Digest of an .ncu-rep: key raw metrics + executed instructions / stall samples per source line (needs ncu on PATH)."""
import csv, subprocess, sys, io
from collections import defaultdict

WANT = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main(path, top=30):
    rows = list(csv.reader(io.StringIO(run(['-i', path, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:75s} {r[i]} {units[i]}")
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    if float(r[i]) > 0.3:
                        print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {float(r[i]):.2f}")
                except ValueError:
                    pass
    rows = list(csv.reader(io.StringIO(run(['-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass']))))
    cur = None
    d = defaultdict(lambda: [0, 0, ''])
    nsass = 0
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] in ('Function Name', 'Line No'):
            continue
        if r[0] == '' and len(r) > 7 and r[7].isdigit():
            nsass += 1
        if r[0].isdigit() and len(r) > 7 and r[7].isdigit():
            k = (cur, int(r[0]))
            d[k][0] += int(r[7]); d[k][1] += int(r[6]) if r[6].isdigit() else 0; d[k][2] = r[1].strip()[:100]
    tot = sum(v[0] for v in d.values()); ts = sum(v[1] for v in d.values()) or 1
    print(f"executed warp-instructions (source-attributed) {tot}, stall samples {ts}, SASS instructions {nsass}")
    for (f, l), v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f:16s}:{l:4d} inst {100 * v[0] / tot:5.1f}% smp {100 * v[1] / ts:5.1f}%  {v[2]}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
