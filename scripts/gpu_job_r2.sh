set -x
for L in 0 8 24; do PF_LOCKSTEP=$L timeout 60 scripts/umma_probe 1000 3000000 300 1 perf 2>&1 | grep -E "rep 4|lockstep|PROBE|candidates|HUNG" ; done > gpurun_out/probe_lockstep.log 2>&1; cat gpurun_out/probe_lockstep.log
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/gputests.log 2>&1; tail -12 gpurun_out/gputests.log
timeout 120 python bench.py --steps 5 --warmup 3 --secondary nominal --no-cpu-baseline --pipe-shape 3 > gpurun_out/bench_shape3.json 2> gpurun_out/bench_shape3.err; echo rc=$?
python - <<'PY'
import json
for S in (0,3):
    try:
        j=json.loads(open(f'gpurun_out/bench_shape{S}.json').read().strip().splitlines()[-1])
    except Exception as ex:
        print("shape",S,"no result",ex); continue
    print("shape",S,"headline",round(j["value"]),"stage",j["roofline"]["stage_ms_per_step"]["pipe"])
    for s in j["config"]["secondary"]:
        print("   ",s["name"][:30],round(s.get("queries_per_s",0)), s.get("stage_ms"))
PY
