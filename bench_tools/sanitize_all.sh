set -u
out=gpurun_out/r02_compute_sanitizer.txt
: > $out
run() { tool=$1; shift; echo "## $tool: $*" >> $out; JU_NO_GRAPH=1 JU_WAIT_TIMEOUT_MS=600000 timeout 900 compute-sanitizer --tool $tool python bench_tools/sanitize_run.py "$@" 2>&1 | grep -E "checksum|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard" | head -12 | sed "s/^/$tool: /" >> $out; }
run memcheck small 3 2
run memcheck psp_fast 2 3
run initcheck tiny 3 2
run racecheck tiny 3 2
run synccheck tiny 3 2
run racecheck small 2 3
cat $out
