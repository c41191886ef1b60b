P='import json,sys; d=json.loads(sys.stdin.read()); print(TAG, round(d["value"]), round(d["e2e"]["value"]), {k: round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items() if v})'
for c in 1024 2048 3072 4096; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --pipe-chunk $c 2>&1 | tail -1 | python -c "TAG='chunk=$c'; $P"
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
