"""Print one line per kernel launch from an ncu --csv log with several metrics."""
import csv
import sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
out, order = {}, []
for r in rows:
    k = (r['ID'], r['Kernel Name'])
    if k not in out:
        out[k] = {}
        order.append(k)
    out[k][r['Metric Name']] = (r['Metric Value'], r['Metric Unit'])
for k in order:
    m = out[k]
    def g(name, scale=1.0, fmt='{:.3f}'):
        v = m.get(name)
        return fmt.format(float(v[0].replace(',', ''))*scale) if v else '-'
    t_unit = m.get('gpu__time_duration.sum', ('0', 'ns'))[1]
    ts = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(t_unit, 1e-6)
    bs = lambda n: {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}.get(m.get(n, ('0', 'byte'))[1], 1e-9)
    print(f"{g('gpu__time_duration.sum', ts):>8} ms  rd {g('dram__bytes_read.sum', bs('dram__bytes_read.sum')):>6} GB  wr {g('dram__bytes_write.sum', bs('dram__bytes_write.sum')):>6} GB  "
          f"inst {g('smsp__inst_executed.sum', 1e-6, '{:.1f}'):>7} M  fp64 {g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 1, '{:.1f}'):>5}%  "
          f"issue {g('smsp__issue_active.avg.pct_of_peak_sustained_active', 1, '{:.1f}'):>5}%  {k[1][:70]}")
