"""Summarise an ncu --set full report: one line of key metrics per captured launch.
usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'dur'), ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs'),
        ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('lts__t_bytes.sum', 'l2_bytes'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%active'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor%elapsed'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'),
        ('sm__cycles_elapsed.max', 'cycles')]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[idx['Kernel Name']][:48]
        parts = []
        for key, label in WANT:
            if key in idx:
                parts.append(f'{label}={r[idx[key]]}{units[idx[key]] if units[idx[key]] not in ("", "%") else ""}')
        print(name, ' '.join(parts))


if __name__ == '__main__':
    main(sys.argv[1])
