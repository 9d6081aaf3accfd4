#!/bin/bash
# full GPU test suite with the new model kernels in the library; neural renderer end to end at 1080p + its ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method thread --durations=8 > gpurun_out/pytest_gpu_aj.log 2>&1; echo "pytest rc=$?"; tail -18 gpurun_out/pytest_gpu_aj.log
timeout 600 python tools/bench_disney_render.py > gpurun_out/disney_render_aj.log 2>&1; echo "render rc=$?"; cut -c1-500 gpurun_out/disney_render_aj.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_disney_aj.csv python tools/bench_disney_render.py 1920 1080 512 fast > gpurun_out/ncu_disney_aj.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_disney_aj.csv')) if len(r) > 10 and r[0].isdigit()]
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.reader(open('gpurun_out/launches_disney_aj.csv')):
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: v = float(d['Metric Value'].replace(',', ''))
        except ValueError: continue
        unit = d['Metric Unit']
        v = v / 1e3 if unit in ('ns', 'nsecond') else v * 1e3 if unit in ('ms', 'msecond') else v
        k = d['Kernel Name'][:60]; agg[k][0] += 1; agg[k][1] += v
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]: print(f'{us/1e3:10.2f} ms {n:6d} x  {k}')
PY
