P='import json,sys; d=json.loads(sys.stdin.read()); print(TAG, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["config"]["index_upload_s"], {k: round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()})'
for cfg in "0 1024 0" "0 1024 256" "1 1024 0" "1 1024 256" "1 2048 256" "1 2048 1024"; do
  set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipeline $1 --pipe-chunk $2 --placement $3 2>&1 | tail -1 | python -c "TAG='pipeline=$1 chunk=$2 placement=$3'; $P"
done
