"""Summarise one kernel of an .ncu-rep (read on the CPU box): headline metrics, executed-instruction mix
by opcode, stall reasons, and the hottest source lines. Usage: ncu_summary.py report.ncu-rep [top_lines]"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    which = sys.argv[3] if len(sys.argv) > 3 else None      # substring of the kernel name (reports with several launches)
    raw = ncu(rep, "raw")
    h = raw[0]
    iK = h.index("Kernel Name")
    row = next((r for r in raw[2:] if which is None or which in r[iK]), raw[2])
    print("# kernel:", row[iK])
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"{k:72s} {row[i]:>16s} {raw[1][i]}")
    src = ncu(rep, "source")
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    s0 = next((i for i in starts if which is None or which in src[i][1]), starts[0])
    s1 = min([i for i in starts if i > s0] + [len(src)])
    hs = src[s0 + 1]
    data = [r for r in src[s0 + 2:s1] if len(r) > hs.index('Instructions Executed')]
    iS, iN, iP = hs.index('Source'), hs.index('Instructions Executed'), hs.index('# Samples')
    byop, samp = Counter(), Counter()
    for r in data:
        t = r[iS].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0].rstrip(';')
        byop[op] += int(r[iN] or 0)
        samp[op] += int(r[iP] or 0)
    tot, ts = sum(byop.values()), max(sum(samp.values()), 1)
    print(f"# instruction mix (warp instructions executed: {tot})")
    for op, n in byop.most_common(top + 6):
        print(f"  {op:10s} {n:>14d} {100*n/tot:5.1f}%   stall samples {100*samp[op]/ts:5.1f}%")
    print("# stall reasons (samples)")
    st = {k: sum(int(r[hs.index(k)] or 0) for r in data) for k in hs if k.startswith('stall_') and 'Not Issued' not in k}
    for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {k:28s} {v:>10d} {100*v/max(sum(st.values()),1):5.1f}%")


if __name__ == "__main__":
    main()
