set -x
for L in 0 4 8 32; do PF_LOCKSTEP=$L timeout 100 scripts/umma_probe 1000 3000000 300 1 perf 2>&1 | grep -E "rep 4|lockstep|PROBE|candidates" ; done > gpurun_out/probe_lockstep.log 2>&1; cat gpurun_out/probe_lockstep.log
timeout 300 python -m pytest tests/test_pipeline_gpu.py tests/test_prefilter_gpu.py tests/test_ivfadc_gpu.py -q --timeout 280 2>&1 | tail -5
for S in 0 3; do timeout 300 python bench.py --steps 5 --warmup 3 --secondary nominal --no-cpu-baseline --pipe-shape $S > gpurun_out/bench_shape$S.json 2> gpurun_out/bench_shape$S.err; done
python - <<'PY'
import json
for S in (0,3):
    j=json.loads(open(f'gpurun_out/bench_shape{S}.json').read().strip().splitlines()[-1])
    print("shape",S,"headline",round(j["value"]),"stage",j["roofline"]["stage_ms_per_step"]["pipe"])
    for s in j["config"]["secondary"]:
        print("   ",s["name"][:30],round(s.get("queries_per_s",0)), s.get("stage_ms"))
PY
