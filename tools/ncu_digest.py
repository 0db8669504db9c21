"""Digest of `ncu --set full` captures (raw page CSV): one block per kernel launch with the metrics the roofline uses.

    python tools/ncu_digest.py gpurun_out/prof_r1b_*.ncu-rep > profiles/ncu_r1b_digest.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (realtime, elapsed)"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu.sum", "smem bank conflicts"),
]


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"# {path}: no launches captured")
            continue
        hdr, units = rows[0], rows[1]
        print(f"# {path}")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print(f"kernel: {d.get('Kernel Name', '?')[:110]}")
            for k, label in KEYS:
                if k in d and d[k] != "":
                    print(f"    {label:36s} {d[k]:>16s} {u.get(k, '')}")
            rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
            if rd and wr:
                print(f"    {'dram traffic (read+write)':36s} {float(rd) + float(wr):16.3f} {u.get('dram__bytes_read.sum', '')}")
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
