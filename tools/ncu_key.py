"""Key metrics of every kernel of an .ncu-rep (ncu --set full), one block per launch.   python tools/ncu_key.py file.ncu-rep [name-filter]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'sm__inst_executed_pipe_tex.sum', 'l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed']
STALL = 'smsp__average_warps_issue_stalled_'


def main():
    rows = list(csv.reader(subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ''
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        if flt not in r[ki]:
            continue
        print('---', r[ki][:90])
        for w in WANT:
            if w in hdr:
                print(f'  {w:75s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}')
        st = [(float(r[i] or 0), hdr[i][len(STALL):-len('_per_issue_active.ratio')]) for i in range(len(hdr)) if hdr[i].startswith(STALL) and hdr[i].endswith('_per_issue_active.ratio')]
        print('  stalls per issue:', ', '.join(f'{n} {v:.2f}' for v, n in sorted(st, reverse=True)[:7]))


if __name__ == '__main__':
    main()
