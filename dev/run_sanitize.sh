set -x
python dev/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1; tail -3 gpurun_out/sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python dev/sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
