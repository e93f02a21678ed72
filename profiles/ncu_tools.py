"""Helpers to read an .ncu-rep here (no GPU): key metrics per kernel and stall samples per SASS line.
    python profiles/ncu_tools.py metrics gpurun_out/prof.ncu-rep
    python profiles/ncu_tools.py hot gpurun_out/prof.ncu-rep <kernel-index> [top]
"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__cycles_elapsed.max', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sectors_op_write.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def metrics(rep):
    hdr, units, rows = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    for k, r in enumerate(rows):
        print('---- [%d] %s' % (k, r[idx['Kernel Name']][:90]))
        for key in KEYS:
            if key in idx:
                print('   %-95s %s %s' % (key, r[idx[key]], units[idx[key]]))


def hot(rep, kidx, top=25):
    hdr, units, rows = raw(rep)
    name_i = hdr.index('Kernel Name')
    # source page of the k-th launch
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(kidx), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = None
    data = []
    for r in rows:
        if 'Source' in r and '# Samples' in r:
            if h is not None:
                break                       # second (duplicate) view
            h = r
            continue
        if h is None or len(r) < len(h):
            continue
        try:
            data.append((int(r[h.index('# Samples')]), int(r[h.index('Instructions Executed')]), r[h.index('Source')].strip()))
        except ValueError:
            pass
    tot = sum(d[0] for d in data)
    print('total samples', tot, 'instructions', sum(d[1] for d in data))
    for k, d in sorted(enumerate(data), key=lambda kd: -kd[1][0])[:top]:
        print('%5d  samples %5d (%.1f%%)  exec %8d  %s' % (k, d[0], 100.0 * d[0] / max(tot, 1), d[1], d[2]))
    return data


if __name__ == '__main__':
    if sys.argv[1] == 'metrics':
        metrics(sys.argv[2])
    else:
        hot(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 25)
