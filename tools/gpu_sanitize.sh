#!/bin/bash
# compute-sanitizer evidence: racecheck, synccheck and memcheck over tools/sanitize_step.py, with the
# key-by-key set insert and with the bulk page build forced on small sets
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  for bulk in 1000000000 1; do
    log=gpurun_out/r02_sanitizer_${tool}_bulk${bulk}.log
    SGPU_IDSET_BULK_MIN=$bulk timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py 3000 > $log 2>&1
    echo "$tool bulk_min=$bulk rc=$? : $(grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_step ok|Error|AssertionError" $log | tr '\n' ' ' | cut -c1-300)"
  done
done
