"""Summarise an .ncu-rep: headline metrics + SASS regions grouped by execution count."""
import csv, subprocess, sys, io

def main(rep, groups=True):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'smsp__thread_inst_executed_per_inst_executed.ratio',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'launch__grid_size', 'launch__block_size', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
            'launch__waves_per_multiprocessor', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_fma.sum',
            'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum']
    for r in rows[2:]:
        for w in want:
            if w in h:
                i = h.index(w)
                print(f"{w:70s} {r[i]:>20s} {rows[1][i]}")
        for i, c in enumerate(h):
            if c.startswith('smsp__average_warps_issue_stalled') and c.endswith('per_issue_active.ratio'):
                try:
                    v = float(r[i])
                    if v > 0.2:
                        print(f"  stall {c[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:.2f}")
                except ValueError:
                    pass
    if not groups:
        return
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    isrc, ie, it, iss = h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[ie]), int(r[it]), int(r[iss]), r[isrc]))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data)
    print('total inst', tot, 'stall samples', sum(d[2] for d in data))
    i = 0
    while i < len(data):
        j, c = i, data[i][0]
        while j < len(data) and abs(data[j][0] - c) <= 0.15 * max(c, 1):
            j += 1
        s = sum(d[0] for d in data[i:j]); st = sum(d[2] for d in data[i:j]); thr = sum(d[1] for d in data[i:j])
        if s > tot * 0.006:
            print(f"[{i:4d}-{j - 1:4d}] n={j - i:3d} exec/inst={c / 1e3:8.0f}k total={s / 1e6:7.2f}M ({100 * s / tot:4.1f}%) lanes={thr / max(s, 1):4.1f} stalls={st:5d}  {data[i][3][:56]}")
        i = j

if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) < 3)
