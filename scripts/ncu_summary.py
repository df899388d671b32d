#!/usr/bin/env python
"""One table row per kernel launch from an `ncu --page raw --csv` export (the format of profiles/*_raw.csv):
duration, DRAM bytes read / written (= `roofline.traffic`), DRAM throughput as a fraction of the measured copy peak,
registers, achieved occupancy, L2 hit rate, tensor-pipe activity, issue-slot utilisation.

  ncu -i gpurun_out/r02_gemm_tf32.ncu-rep --page raw --csv > profiles/r02_gemm_tf32_raw.csv
  python scripts/ncu_summary.py profiles/r02_gemm_tf32_raw.csv [more.csv ...]
Metric names are matched by suffix because ncu prefixes some of them with their section (e.g. `FBSP.TriageCompute.`)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    ('ms', 'gpu__time_duration.sum'),
    ('dram_rd_GB', 'dram__bytes_read.sum'),
    ('dram_wr_GB', 'dram__bytes_write.sum'),
    ('regs', 'launch__registers_per_thread'),
    ('occ_%', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
    ('l2_hit_%', 'lts__t_sector_hit_rate.pct'),
    ('tensor_%', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'),
    ('tensor_%', 'sm__ops_path_tensor_src_tf32_dst_fp32.avg.pct_of_peak_sustained_elapsed'),
    ('issue_%', 'sm__issue_active.avg.pct_of_peak_sustained_active'),
    ('issue_%', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
    ('smem_KB', 'launch__shared_mem_per_block_dynamic'),
]
UNIT_SCALE = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3,
              'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0, 'Tbyte': 1e3}


def column(header, suffix):
    for i, name in enumerate(header):
        if name == suffix or name.endswith('.' + suffix):
            return i
    return None


def peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    return float(json.load(open(path))['hbm_gbs']) if os.path.isfile(path) else 6650.0


def summarize(path):
    rows = list(csv.reader(open(path, errors='replace')))
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')          # ncu may print banner lines first
    header, units, launches = rows[start], rows[start + 1], rows[start + 2:]
    name_col = column(header, 'Kernel Name')
    out = []
    for r in launches:
        if len(r) != len(header):
            continue
        rec = {'kernel': r[name_col].split('(')[0].replace('void ', '')[:70]}
        for key, metric in WANT:
            i = column(header, metric)
            if i is None or key in rec or not r[i]:
                continue
            try:
                val = float(r[i].replace(',', ''))
            except ValueError:
                continue
            scale = UNIT_SCALE.get(units[i], 1.0) if key in ('ms', 'dram_rd_GB', 'dram_wr_GB') else 1.0
            rec[key] = val * scale if key != 'smem_KB' else val * (1e-3 if units[i] == 'byte/block' else 1.0)
        if 'ms' in rec and 'dram_rd_GB' in rec:
            traffic = rec['dram_rd_GB'] + rec.get('dram_wr_GB', 0.0)
            rec['traffic_GB'] = traffic
            rec['dram_TB/s'] = traffic / rec['ms']
            rec['of_copy_peak'] = traffic / rec['ms'] * 1e3 / peak_gbs()
        out.append(rec)
    return out


def main():
    keys = ['kernel', 'ms', 'traffic_GB', 'dram_TB/s', 'of_copy_peak', 'regs', 'occ_%', 'l2_hit_%', 'tensor_%', 'issue_%', 'smem_KB']
    print('| file | ' + ' | '.join(keys) + ' |')
    print('|' + '---|' * (len(keys) + 1))
    for path in sys.argv[1:]:
        for rec in summarize(path):
            cells = [f'{rec[k]:.3f}' if isinstance(rec.get(k), float) else str(rec.get(k, '')) for k in keys]
            print(f'| {os.path.basename(path)} | ' + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    main()
