cd /root/repo
for lg in 20 21; do
  echo "=== HT_LOG2=$lg"; FGNN_DIAG_HT_LOG2=$lg timeout 300 python tools/bench_chain.py --ks 1,4 2>&1 | grep "^K="
done
echo "=== versioned off"; FGNN_HT_VERSIONED=0 timeout 300 python tools/bench_chain.py --ks 1,4 2>&1 | grep "^K="
echo "=== GCN uk"; timeout 600 python tools/bench_chain.py --workload uk-2006-05 --fanout 5,10,15 --ks 1,4 --reps 10 2>&1 | grep "^K="
echo "=== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "random_walk_topk or factored_training_example or arch5_forked" 2>&1 | tail -15
