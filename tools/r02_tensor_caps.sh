O=gpurun_out/r02_tensor; mkdir -p $O
for cfg in "ml20m_item_cosine_k40 exact cos" "ml20m_item_msd_k40 exact msd" "ml20m_item_pearson_k40 sums pearson_sums"; do
  set -- $cfg
  python bench.py --workload $1 --sim-path tensor --pearson-mode $2 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_$3.jsonl 2> $O/bench_$3.err && \
  ncu --set full --clock-control none -k regex:sim_tensor -s 3 -c 1 -o $O/r02_tensor_$3 -f python bench.py --workload $1 --sim-path tensor --pearson-mode $2 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_$3.log 2>&1
done
ls -la $O
