#!/bin/bash
# launch list of the 2-D eager step (ncu per-launch durations; cold-cache, serialised: shares only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02u_launches_2d.csv \
  python tools/bench_2d.py --eager --skip_torch --steps 1 --warmup 1 > gpurun_out/r02u_ncu_2d.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r02u_ncu_2d.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = []
with open('gpurun_out/r02u_launches_2d.csv') as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
allrows = list(r)
# keep the launches of the LAST complete step: use the last 1/3 of the rows as an approximation is fragile;
# instead aggregate everything and report shares
for row in allrows:
    name = row['Kernel Name'].split('(')[0].split('<')[0]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    ms = v / 1e6 if unit in ('ns', 'nsecond') else (v / 1e3 if unit in ('us', 'usecond') else v)
    agg[name][0] += 1; agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"{len(allrows)} launches, {tot:.2f} ms total (all captured steps)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{k[:70]:70s} n={v[0]:5d} {v[1]:8.3f} ms {100*v[1]/tot:5.1f} %")
PY
