P='import json,sys; d=json.loads(sys.stdin.read()); print(TAG, round(d["value"]), d["config"]["index_upload_s"], {k: round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items() if v})'
for pl in 64 256 1024 4096; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --placement $pl 2>&1 | tail -1 | python -c "TAG='placement=$pl'; $P"
done
