P='import json,sys; d=json.loads(sys.stdin.read()); print(TAG, round(d["value"]), round(d["e2e"]["value"]), d["gpu_launches"], {k: round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items() if v})'
timeout 200 python -m pytest tests/test_pipeline_gpu.py -x -q 2>&1 | tail -2
for cfg in "0 2048" "1 2048" "1 1024" "1 1536" "1 3072"; do
  set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipe-shape 5 --pipe-ramp $1 --pipe-chunk $2 2>&1 | tail -1 | python -c "TAG='ramp=$1 chunk=$2'; $P"
done
