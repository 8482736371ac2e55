"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of numbers DESIGN.md / bench.py cite."""
import csv
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_uniform.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for k in KEYS:
            if k in hdr:
                d[k] = r[hdr.index(k)] + ' ' + units[hdr.index(k)]
        for i, h in enumerate(hdr):
            if (('pipe_tensor' in h and 'pct_of_peak_sustained_active' in h) or ('warp_issue_stalled' in h and 'per_warp_active.pct' in h)) and h not in d:
                try:
                    if float(r[i]) > 2:
                        d[h] = r[i] + ' ' + units[i]
                except ValueError:
                    pass
        res.append(d)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
