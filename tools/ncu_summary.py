#!/usr/bin/env python
"""Condense an `ncu --set full` capture into the text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/row_solve_tc_raw.csv [gpurun_out/row_solve_tc.ncu-rep] > profiles/rNN_..._ncu_summary.txt

The raw CSV is `ncu -i X.ncu-rep --page raw --csv`; with the report itself as second argument the per-instruction
stall samples of every launch are added (`--page source --csv`, needs ncu on PATH)."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for k, d in enumerate(data):
        print(f"=== launch {k}: {d[idx['Kernel Name']]}")
        for key, label in KEYS:
            if key in idx:
                print(f"  {label:32s} {d[idx[key]]} {units[idx[key]]}")
        st = [(h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], float(d[idx[h]]))
              for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
        st.sort(key=lambda x: -x[1])
        print("  warp stalls per issue: " + ", ".join(f"{n}={v:.2f}" for n, v in st[:8]))
    if len(sys.argv) > 2:
        for k in range(len(data)):
            out = subprocess.run(["ncu", "-i", sys.argv[2], "--page", "source", "--csv", "--launch-skip", str(k),
                                  "--launch-count", "1"], capture_output=True, text=True).stdout
            r = list(csv.reader(out.splitlines()))
            if len(r) < 3:
                continue
            h = r[1]
            ix = {x: i for i, x in enumerate(h)}
            body = [x for x in r[2:] if len(x) >= len(h) - 2 and x[ix['# Samples']].strip().isdigit()][::2]
            tot = sum(int(x[ix['# Samples']]) for x in body) or 1
            reasons = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
            agg = sorted(((c, sum(int(x[ix[c]] or 0) for x in body)) for c in reasons), key=lambda t: -t[1])
            print(f"=== launch {k}: stall samples {tot}: " + ", ".join(f"{c[6:]}={100 * v / tot:.0f}%" for c, v in agg[:8]))
            for x in sorted(body, key=lambda x: -int(x[ix['# Samples']]))[:14]:
                st = {c: int(x[ix[c]] or 0) for c in reasons}
                m = max(st, key=st.get)
                print(f"    {100 * int(x[ix['# Samples']]) / tot:5.1f}%  {x[ix['Source']][:86]:86s} {m[6:]}")


if __name__ == "__main__":
    main()
