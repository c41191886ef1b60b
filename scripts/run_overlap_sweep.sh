timeout 600 python -m pytest tests/test_ivfadc_gpu.py -q -k chunks 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read()); print(TAG, round(d["value"]), round(d["e2e"]["value"]), {k: round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()})'
for cfg in "0 0 512 2048" "1 0 512 2048" "1 1 512 2048" "1 1 512 1024" "1 1 256 1024" "1 2 256 1024" "1 1 1024 1024" "1 1 512 512"; do
  set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --overlap $1 --lut-ctas $2 --lut-tile $3 --chunk $4 2>&1 | tail -1 | python -c "TAG='ov=$1 ctas=$2 tile=$3 chunk=$4'; $P"
done
