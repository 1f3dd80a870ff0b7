mkdir -p gpurun_out
{ echo "compute-sanitizer, tools/sanitize.py (every kernel incl. the round-2 update pipeline at full map size, append, cull):";
for tool in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | grep -E "COMPUTE-SANITIZER|found|big state|batched|ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK|hazard|Error|error" | head -30; echo "$tool rc=$?"; done; } > gpurun_out/r02_compute_sanitizer.log 2>&1
cat gpurun_out/r02_compute_sanitizer.log
