set -x
timeout 600 python -m pytest tests/test_build_gpu.py tests/test_ivfadc_gpu.py tests/test_append_gpu.py -m gpu -q --timeout 300 2>&1 | tail -4
timeout 600 python bench.py --secondary 3 --no-cpu-baseline --steps 3 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; tail -c 300 gpurun_out/bench_tmp.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_tmp.json').read().strip().splitlines()[-1])
print("headline", round(j["value"]))
for s in j["config"]["secondary"]:
    print("   ", s.get("name","")[:40], s.get("seconds"), s.get("stage_ms"), s.get("roofline",{}).get("frac"), s.get("equals_reference_on_sample",{}).get("ok"), s.get("error"))
PY
FB_TRACE_BUILD=1 timeout 300 python scripts/bench_upload.py > gpurun_out/r2_upload.json 2> gpurun_out/upload.err; cat gpurun_out/r2_upload.json; grep "fb build" gpurun_out/upload.err | sed -n 1,6p
