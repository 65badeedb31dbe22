#!/bin/bash
# One parametrised GPU-box script (replaces round 1's per-call scratch scripts).
#   gpurun [--gpus N] --timeout T -- 'bash tools/gpu_run.sh <tag> <stage> [<stage> ...]'
# Stages write into gpurun_out/<tag>_*; copy what should be judged into profiles/.
#   tests:<expr>     pytest -m gpu -k <expr>        (expr "all" = whole GPU suite)
#   bench[:args]     python bench.py <args>         (N = 1)
#   benchN:<n>[:args] torchrun --nproc-per-node n bench.py --gpus n <args>
#   ncu-list[:args]  launch list (gpu__time_duration) of bench.py <args>
#   ncu-full:<regex>[:args]  ncu --set full of the kernels matching <regex> in bench.py <args>
#   py:<script args> python <script args>
set -u
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
for stage in "$@"; do
  kind=${stage%%:*}; rest=${stage#*:}; [ "$kind" == "$stage" ] && rest=""
  echo "=== [$tag] $stage"
  case $kind in
    tests)
      expr=$rest
      if [ "$expr" == "all" ]; then
        timeout 2400 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -25 | tee gpurun_out/${tag}_tests_all.log
      else
        timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -k "$expr" 2>&1 | tail -60 | tee gpurun_out/${tag}_tests.log
      fi ;;
    bench)
      timeout 1200 python bench.py $rest > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      echo "rc=$?"; tail -c 1500 gpurun_out/${tag}_bench.err; head -c 6000 gpurun_out/${tag}_bench.json ;;
    benchN)
      n=${rest%%:*}; args=${rest#*:}; [ "$n" == "$rest" ] && args=""
      timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $n $args > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
      echo "rc=$?"; tail -c 1500 gpurun_out/${tag}_bench_n$n.err; head -c 6000 gpurun_out/${tag}_bench_n$n.json ;;
    ncu-list)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
        --log-file gpurun_out/${tag}_launches.csv python bench.py $rest > gpurun_out/${tag}_ncu_list.log 2>&1
      echo "rc=$?"; tail -3 gpurun_out/${tag}_ncu_list.log | head -c 600 ;;
    ncu-full)
      rx=${rest%%:*}; args=${rest#*:}; [ "$rx" == "$rest" ] && args=""
      timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$rx" -c 12 \
        -o gpurun_out/${tag}_full -f python bench.py $args > gpurun_out/${tag}_ncu_full.log 2>&1
      echo "rc=$?"; tail -3 gpurun_out/${tag}_ncu_full.log | head -c 600 ;;
    py)
      timeout 1500 python $rest 2>&1 | tail -40 | tee gpurun_out/${tag}_py.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done
