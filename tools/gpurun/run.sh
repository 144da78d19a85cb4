#!/bin/bash
# run.sh -- the ONE script behind every gpurun call of this repo (replaces the per-call scripts of round 1).
#   bash tools/gpurun/run.sh tests [pytest args]          GPU test suite (default: tests -m gpu -q)
#   bash tools/gpurun/run.sh time  CASE...                event-timed GB/s: N:transform:direction:ordered[:d]  (tools/time_cases.py)
#   bash tools/gpurun/run.sh bench [bench.py args]        bench.py (JSON line -> gpurun_out/bench.json)
#   bash tools/gpurun/run.sh ncu   NAME REGEX N tr log2batch dir   one `ncu --set full` capture -> gpurun_out/NAME.ncu-rep
#   bash tools/gpurun/run.sh launches                     launch list of `bench.py --steps 2 --warmup 1` -> gpurun_out/launches_bench.csv
#   bash tools/gpurun/run.sh sanitize                     compute-sanitizer memcheck + racecheck over tools/sanitize_cases.py
# Environment variables (PFFFT_B200_*) pass through.  Several modes can be chained with `--`.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run_one() {
  mode=$1; shift
  case "$mode" in
    tests)    if [ $# -eq 0 ]; then set -- tests -m gpu -q; fi
              timeout 2400 python -m pytest "$@" 2>&1 | tail -n 15 ;;
    time)     timeout 900 python tools/time_cases.py "$@" ;;
    bench)    timeout 1200 python bench.py "$@" | tee gpurun_out/bench.json ;;
    ncu)      name=$1; regex=$2; shift 2
              timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s 2 -c 1 -f -o "gpurun_out/$name" \
                python tools/prof_case.py "$@" > "gpurun_out/$name.log" 2>&1; tail -n 2 "gpurun_out/$name.log" ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
                python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; tail -n 3 gpurun_out/launches_bench.csv ;;
    sanitize) for tool in memcheck racecheck; do
                timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_cases.py > "gpurun_out/sanitize_$tool.log" 2>&1
                echo "$tool rc=$?"; tail -n 4 "gpurun_out/sanitize_$tool.log"; done ;;
    *) echo "run.sh: unknown mode $mode"; return 2 ;;
  esac
}
args=()
for a in "$@"; do
  if [ "$a" == "--" ]; then echo "== ${args[*]}"; run_one "${args[@]}"; args=(); else args+=("$a"); fi
done
if [ ${#args[@]} -gt 0 ]; then echo "== ${args[*]}"; run_one "${args[@]}"; fi
