set -x
timeout 300 python bench.py --secondary none --no-cpu-baseline --batch 262144 --steps 3 --warmup 3 > gpurun_out/bench_8shards.json 2> gpurun_out/bench_8shards.err; tail -c 300 gpurun_out/bench_8shards.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_8shards.json').read().strip().splitlines()[-1])
print("8 shards on one GPU: value", round(j["value"]), j["config"]["exact_path_queries_per_step"], j["config"]["exact_path_reasons_per_step"], j["config"]["per_rank"][0]["stage_ms_per_step"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:'ivfadc|coarse_|count_rows|adc_scan|finalize|lut_|prefilter|pf_|subset_|join_|ivpq_|batch_unfilled|analogy|knn_|exact|rerank|cosine|table_|place_rows|pack_rows' --log-file gpurun_out/r2_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -c 200 gpurun_out/ncu_bench.log; wc -l gpurun_out/r2_launch_list.csv
