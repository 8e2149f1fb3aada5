#!/bin/bash
# round 2, session 58 (1 GPU): compute-sanitizer memcheck + racecheck over the module's kernels on small scenes
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/sanitize_small.py > $OUT/s58_plain.log 2>&1; tail -2 $OUT/s58_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > $OUT/s58_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|sanitize_small OK|Invalid|out of bounds" $OUT/s58_memcheck.log | head -5
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py > $OUT/s58_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|sanitize_small OK|hazard" $OUT/s58_racecheck.log | head -8
