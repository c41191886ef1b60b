P='import json,sys; d=json.loads(sys.stdin.read()); print(TAG, round(d["value"]), round(d["e2e"]["value"]), {k: round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items() if v})'
timeout 200 python -m pytest tests/test_pipeline_gpu.py -x -q 2>&1 | tail -2
for shape in ${SHAPES:-0 5 6 7}; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipe-chunk ${CHUNK:-2048} --pipe-shape $shape 2>&1 | tail -1 | python -c "TAG='shape=$shape'; $P"
done
