#!/bin/bash
# The GPU runs behind profiles/r02_* (each line is one `gpurun` call; see profiles/README.md).  Usage: tools/profile_round.sh <name>
set -e
G="tools/gpu.sh"
CMD="python bench.py --timepoints 2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs"
case "$1" in
  tests)    $G 900 'python -m pytest tests -m gpu -q 2>&1 | tail -4' ;;
  tests2)   $G --gpus 2 900 'python -m pytest tests -m gpu -q 2>&1 | tail -4' ;;
  default)  $G 1200 'python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 400 gpurun_out/bench_default.log' ;;
  ncu)      $G 1200 "$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fp64.csv $CMD > gpurun_out/ncu_a.log 2>&1; ncu --set full --clock-control none --import-source on --launch-skip 13 -c 6 -o gpurun_out/prof_all -f $CMD > gpurun_out/ncu_c.log 2>&1" ;;
  cfg5)     $G --gpus "${2:-2}" 900 "python bench.py --gpus ${2:-2} --workload cfg5 --steps 2 --warmup 2 > gpurun_out/bench_cfg5_${2:-2}gpu.log 2>&1; tail -c 400 gpurun_out/bench_cfg5_${2:-2}gpu.log" ;;
  scale)    $G --gpus "${2:-8}" 900 "python bench.py --gpus ${2:-8} --steps 2 --warmup 3 > gpurun_out/bench_${2:-8}gpu.log 2>&1; tail -c 400 gpurun_out/bench_${2:-8}gpu.log" ;;
  pcie)     $G --gpus 4 600 'python tools/pcie_bw.py --gpus 4; python tools/pcie_bw.py --gpus 4 --bind' ;;
  *) echo "usage: $0 tests|tests2|default|ncu|cfg5 [N]|scale [N]|pcie"; exit 1 ;;
esac
# afterwards, here (no GPU):  python tools/ncu_list.py gpurun_out/launches_fp64.csv
#                             python tools/ncu_traffic.py cfg4_fp64=gpurun_out/launches_fp64.csv:134217728:13:6 > profiles/rNN_traffic.json
#                             ncu -i gpurun_out/prof_all.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv "# header" > profiles/rNN_all_kernels_ncu_full.txt
