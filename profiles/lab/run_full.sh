#!/bin/bash
# full GPU validation: smoke, the whole -m gpu suite, the default bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench_err.txt; tail -3 gpurun_out/bench_err.txt
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_line.json').read().strip().splitlines()[-1])
print('value', r['value'], 'ms', r['ms_per_step'], 'e2e', r['e2e']['value'], r['e2e']['ms_per_step'])
print('roofline', {k: r['roofline'][k] for k in ('achieved', 'peak', 'frac', 'kernel_ms', 'call_ms', 'launches_per_call', 'frac_of_call')})
print('variants', r['kernel_variants'])
for k, v in (r.get('workloads') or {}).items():
    print(k, {a: v.get(a) for a in ('kernel_ms', 'call_ms', 'device_path_ms_per_step', 'e2e_ms_per_step')}, v.get('roofline'))
print('other', {k: r['other_precision'][k] for k in ('kernel_ms', 'device_path_ms_per_step')})
print('clocks', r['clocks'])
print('cpu', r['cpu_baseline'])
PY
} > gpurun_out/full_lab.txt 2>&1
cat gpurun_out/full_lab.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python - <<'PY' >> gpurun_out/full_lab.txt
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/final_launches_bench.csv')) if len(r) > 10 and r[0].isdigit()]
tot = collections.Counter(); n = collections.Counter()
for r in rows:
    name = r[4].split('(')[0][:70]; tot[name] += float(r[-1]); n[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(12): print(f'{v / s * 100:6.2f} %  {n[k]:4d} x  {k}')
PY
tail -14 gpurun_out/full_lab.txt
