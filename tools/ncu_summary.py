#!/usr/bin/env python
"""Condense one kernel of an .ncu-rep (ncu --set full) into the JSON summary kept under profiles/.
  python tools/ncu_summary.py report.ncu-rep --units N --alg-bytes B --capture "..." > profiles/x.json"""
import argparse
import csv
import io
import json
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--units', type=float, default=None, help='work units (evaluations / factors) in the captured launch')
    ap.add_argument('--alg-bytes', type=float, default=None, help='algorithmic DRAM bytes of the captured launch')
    ap.add_argument('--capture', default='')
    a = ap.parse_args()
    out = subprocess.run(['ncu', '-i', a.report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, vals = rows[0], rows[2]
    m = dict(zip(hdr, vals))

    def f(k):
        try:
            return float(m[k].replace(',', ''))
        except Exception:
            return None
    rd, wr = f('dram__bytes_read.sum'), f('dram__bytes_write.sum')
    unit = dict(zip(hdr, rows[1]))
    def to_bytes(k, v):
        u = unit.get(k, '')
        return None if v is None else v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(u, 1)
    dur = f('gpu__time_duration.sum')
    dur_ms = None if dur is None else dur * {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}.get(unit.get('gpu__time_duration.sum', 'ns'), 1e-6)
    s = {
        'kernel': m.get('Kernel Name'), 'grid': m.get('Grid Size'), 'block': m.get('Block Size'), 'duration_ms': dur_ms,
        'registers_per_thread': f('launch__registers_per_thread'),
        'dram_bytes_read': to_bytes('dram__bytes_read.sum', rd), 'dram_bytes_write': to_bytes('dram__bytes_write.sum', wr),
        'pipe_fp64_pct': f('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
        'pipe_tensor_dmma_pct': f('sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active') or f('sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active'),
        'sm_throughput_pct': f('sm__throughput.avg.pct_of_peak_sustained_elapsed'),
        'lsu_shared_wavefronts_pct': f('l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed'),
        'warps_active_pct': f('sm__warps_active.avg.pct_of_peak_sustained_active'),
        'issue_active_pct': f('sm__inst_issued.avg.pct_of_peak_sustained_active') or f('smsp__issue_active.avg.pct_of_peak_sustained_active'),
        'occupancy_limit_registers_blocks': f('launch__occupancy_limit_registers'),
        'occupancy_limit_smem_blocks': f('launch__occupancy_limit_shared_mem'),
        'icc_hit_rate_pct': f('sm__icc_request_hit_rate.pct'),
    }
    for k, v in m.items():
        if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio'):
            name = k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]
            try:
                x = float(v)
            except ValueError:
                continue
            if x >= 0.1:
                s['stall_%s_per_issue' % name] = x
    s['capture'] = a.capture
    if s['dram_bytes_read'] is not None:
        s['dram_bytes_per_launch'] = (s['dram_bytes_read'] or 0) + (s['dram_bytes_write'] or 0)
    if a.alg_bytes is not None:
        s['algorithmic_bytes_per_launch'] = a.alg_bytes
    if a.units is not None:
        s['units_in_capture'] = a.units
    print(json.dumps(s, indent=1))


if __name__ == '__main__':
    main()
