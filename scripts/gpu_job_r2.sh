set -x
timeout 600 python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; tail -c 300 gpurun_out/bench_r2_default.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench_r2_default.json').read().strip().splitlines()[-1])
    print("value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "frac", j["roofline"]["frac"], "cpu", j["cpu_baseline"]["value"], j["clocks"], "launches", j["gpu_launches"])
    for s in j["config"]["secondary"]:
        print("   ", s.get("name","")[:60], s.get("seconds"), s.get("queries_per_s"), s.get("roofline",{}).get("frac"), s.get("stage_ms", s.get("stage_ms_rank0")), s.get("equals_reference_on_sample",{}).get("ok"), s.get("error"))
except Exception as ex:
    print("no result", ex)
PY
