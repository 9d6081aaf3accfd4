#!/bin/bash
# One parameterised GPU-box script (replaces the per-session tools/gpu_round1_*.sh files).  Usage under gpurun:
#   gpurun --timeout 1500 -- 'bash tools/gpu_run.sh <tag> <step> [<step> ...]'
# Steps: smoke | tests | tests:<pytest -k expr> | bench | bench:<extra args> | launches | ncu:<kernel regex> | l2 | cmd:<shell command>
# Every step logs to gpurun_out/<tag>_<step>.log and prints a short tail, so a cut-off call can still be read afterwards.
tag=$1; shift
mkdir -p gpurun_out
i=0
for step in "$@"; do
  name=${step%%:*}; arg=""; [[ "$step" == *:* ]] && arg=${step#*:}
  i=$((i + 1)); lname=$name; [ -e gpurun_out/${tag}_${name}.log ] && lname=${name}$i   # a step that appears twice keeps both logs
  log=gpurun_out/${tag}_${lname}.log
  case $name in
    smoke)    timeout 600 python __graft_entry__.py smoke > $log 2>&1; echo "[$step] rc=$?"; tail -2 $log ;;
    tests)    if [ -n "$arg" ]; then timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --timeout-method thread --durations=8 -k "$arg" > $log 2>&1;
              else timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --timeout-method thread --durations=8 > $log 2>&1; fi; echo "[$step] rc=$?"; tail -15 $log ;;
    bench)    timeout 900 python bench.py $arg > $log 2>&1; echo "[$step] rc=$?"; grep '^{' $log | tail -1 > gpurun_out/${tag}_${lname}.json; tail -c 1800 gpurun_out/${tag}_${lname}.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 $arg > $log 2>&1; echo "[$step] rc=$?"; python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_summary.txt; head -12 gpurun_out/${tag}_launch_summary.txt ;;
    ncu)      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$arg -s 2 -c 1 -f -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 1 > $log 2>&1; echo "[$step] rc=$?";
              ncu -i gpurun_out/${tag}_prof.ncu-rep --page details > gpurun_out/${tag}_ncu_details.txt 2>/dev/null; grep -E "Duration|Theoretical Occupancy|Achieved Occupancy|Issue Slots Busy|L2 Hit Rate|L1/TEX Hit Rate|Avg. Active Threads|DRAM Throughput|bank conflicts" gpurun_out/${tag}_ncu_details.txt | head -14 ;;
    l2)       ./tools/microbench/l2_gather > gpurun_out/${tag}_l2_peaks.json 2> $log; echo "[$step] rc=$?"; cat gpurun_out/${tag}_l2_peaks.json ;;
    cmd)      timeout 1500 bash -c "$arg" > $log 2>&1; echo "[$step] rc=$?"; tail -12 $log ;;
    *)        echo "unknown step $step" ;;
  esac
done
