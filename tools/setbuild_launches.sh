#!/bin/bash
# ncu launch list of the set build for (a) the C4-like inline id list, (b) 5 M Illumina-style long ids
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file gpurun_out/l_inline.csv python tools/c4_probe.py --steps 1 --pairs 1000000 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"idset_|lines_tile|scan_" -c 40 --csv --log-file gpurun_out/l_long.csv python tools/long_ids.py --pairs 5000000 --steps 1 > /dev/null 2>&1
python - <<PY
import csv
for f in ("l_inline","l_long"):
    rows=[r for r in csv.reader(open(f"gpurun_out/{f}.csv")) if len(r)>5 and r[0].isdigit()]
    print(f)
    for r in rows[:24]:
        if "scan_" in r[4] and float(r[-1]) < 20000: continue
        print("  ", r[4][:46], r[-1])
PY
