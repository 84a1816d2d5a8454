"""Turn .ncu-rep captures (ncu --set full) into the committed text summary and the per-launch DRAM traffic JSON that
bench.py reads for roofline.traffic.   python tools/ncu_summary.py out.txt out_traffic.json rep1.ncu-rep [rep2 ...]"""
import csv, io, json, subprocess, sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct']
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main(out_txt, out_json, reps):
    traffic = []
    with open(out_txt, 'w') as f:
        f.write("# ncu --set full --clock-control none --import-source on; one block per captured launch (tools/ncu_summary.py)\n")
        for rep in reps:
            raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(raw)))
            if len(rows) < 3:
                continue
            hdr, units = rows[0], rows[1]
            f.write(f"## {rep.split('/')[-1]}\n")
            for r in rows[2:]:
                f.write("---\n")
                rec = {}
                for k in KEYS:
                    if k in hdr:
                        i = hdr.index(k)
                        f.write(f"{k:85s} {r[i]} {units[i]}\n")
                        rec[k] = (r[i], units[i])
                for i, h in enumerate(hdr):
                    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                        try:
                            if float(r[i]) >= 0.5:
                                f.write(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {float(r[i]):.2f}\n")
                        except ValueError:
                            pass
                try:
                    rd = float(rec['dram__bytes_read.sum'][0]) * UNIT[rec['dram__bytes_read.sum'][1]]
                    wr = float(rec['dram__bytes_write.sum'][0]) * UNIT[rec['dram__bytes_write.sum'][1]]
                    us = float(rec['gpu__time_duration.sum'][0]) * {'us': 1.0, 'ms': 1e3, 'ns': 1e-3}.get(rec['gpu__time_duration.sum'][1], 1.0)
                    traffic.append({"kernel": rec['Kernel Name'][0], "grid": rec['Grid Size'][0], "dram_bytes": rd + wr, "us": us})
                except (KeyError, ValueError):
                    pass
    with open(out_json, 'w') as f:
        json.dump(traffic, f, indent=1)
    print(f"{len(traffic)} launches -> {out_txt}, {out_json}")


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
