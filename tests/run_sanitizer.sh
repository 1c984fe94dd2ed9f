#!/bin/bash
# compute-sanitizer over the operator battery (tests/gpu_diag.py), one process per (tool, group) so that a trapped
# kernel cannot poison the next group.  SURVEY.md section 5 asks for a clean memcheck / racecheck run per kernel.
#
#   tests/run_sanitizer.sh <outdir> <tool> <group> [<group> ...]      tool: memcheck | racecheck | synccheck | initcheck
#
# Writes <outdir>/san_<tool>_<group>.log (full log) and appends one line per group to <outdir>/san_summary.txt:
#   <tool> <group> rc=<exit code> errors=<ERROR SUMMARY count> battery="<N passed, M failed>" seconds=<wall>
# The mbarrier watchdog of ptx.cuh (trap after 4e9 clocks) stays armed: a protocol bug under the sanitizer's different
# timing becomes an error in the log, not a hung box.
set -u
out=$1; tool=$2; shift 2
mkdir -p "$out"
cd "$(dirname "$0")/.."
for g in "$@"; do
  log="$out/san_${tool}_${g}.log"
  t0=$(date +%s)
  timeout "${SAN_TIMEOUT:-420}" compute-sanitizer --tool "$tool" --print-limit 30 --error-exitcode 86 \
      python tests/gpu_diag.py "$g" > "$log" 2>&1
  rc=$?
  t1=$(date +%s)
  errs=$(grep -o "ERROR SUMMARY: [0-9]* error" "$log" | head -1 | grep -o "[0-9]*")
  hazards=$(grep -o "RACECHECK SUMMARY: [0-9]* hazard" "$log" | head -1 | grep -o "[0-9]*")
  batt=$(grep -o "===== [0-9]* passed, [0-9]* failed" "$log" | tail -1 | sed 's/===== //')
  echo "$tool $g rc=$rc errors=${errs:-?} hazards=${hazards:--} battery=\"${batt:-none}\" seconds=$((t1 - t0))" | tee -a "$out/san_summary.txt"
done
