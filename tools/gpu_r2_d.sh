#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/c4_probe.py 2>&1 | tail -1
SGPU_IDSET_BULK_MIN=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "golden or paf or txt or reads or sam or bam or diff or idset or c1_ or c3_" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_c4probe.csv python tools/c4_probe.py --steps 1 --pairs 2000000 > gpurun_out/c4probe_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_c4probe.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[:9]: print(r[4][:50], r[-1])
PY
timeout 900 python bench.py --steps 5 > gpurun_out/bench_c4_n1.log 2> gpurun_out/bench_c4_n1.err; tail -3 gpurun_out/bench_c4_n1.err; cat gpurun_out/bench_c4_n1.log
