#!/bin/bash
# round 2, session 66 (1 GPU): compute-sanitizer racecheck over the shared-memory users on the smallest workload (tools/sanitize_small.py --quick)
OUT=gpurun_out; mkdir -p $OUT
export NGI_TRACE_GRID_PCT=6 NGI_LANES=1
timeout 60 python tools/sanitize_small.py --quick > $OUT/s66_plain.log 2>&1; tail -1 $OUT/s66_plain.log
timeout 270 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py --quick > $OUT/s66_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|ok$|OK$|hazard|Error" $OUT/s66_racecheck.log | head -12
