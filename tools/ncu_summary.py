"""One paragraph per kernel from an `ncu --set full` report (development aid):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [...] > profiles/rNN_ncu_full_summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = [('dur us', 'gpu__time_duration.sum', 1e-3), ('dram rd GB', 'dram__bytes_read.sum', None), ('dram wr GB', 'dram__bytes_write.sum', None),
        ('dram %', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 1), ('issue %', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 1),
        ('fp64 pipe %', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 1), ('warps active %', 'sm__warps_active.avg.pct_of_peak_sustained_active', 1),
        ('regs', 'launch__registers_per_thread', 1)]
STALLS = ['barrier', 'long_scoreboard', 'short_scoreboard', 'lg_throttle', 'mio_throttle', 'math_pipe_throttle', 'wait', 'membar', 'not_selected']
UNIT = {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0, 'ns': 1.0, 'us': 1e3, 'ms': 1e6}


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            def val(name):
                i = col.get(name)
                if i is None or r[i] == '':
                    return None
                return float(r[i].replace(',', '')), units[i]
            print(r[col['Kernel Name']][:110])
            parts = []
            for label, name, scale in WANT:
                v = val(name)
                if v is None:
                    continue
                x = v[0]*UNIT.get(v[1], 1.0)*(scale if scale else 1.0) if v[1] in UNIT else v[0]*(scale or 1)
                parts.append(f'{label} {x:.4g}')
            wf, cyc = val('l1tex__data_pipe_lsu_wavefronts.sum'), val('sm__cycles_elapsed.avg')
            nsm = 148
            if wf and cyc:
                parts.append(f'L1TEX data-pipe wavefronts/SM {wf[0]/nsm:.4g}  SM cycles {cyc[0]:.4g}  L1TEX busy {100*wf[0]/nsm/cyc[0]:.0f}%')
            print('    ' + '  '.join(parts))
            st = []
            for s in STALLS:
                v = val(f'smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio')
                if v:
                    st.append(f'{s} {v[0]:.2f}')
            print('    warps stalled per issued instruction: ' + ', '.join(st))


if __name__ == '__main__':
    main()
