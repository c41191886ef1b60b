set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err; tail -c 400 gpurun_out/bench_r2_n4.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench_r2_n4.json').read().strip().splitlines()[-1])
    print("N=4 value", round(j["value"]), "e2e", round(j["e2e"]["value"]), j["clocks"])
    for r in j["config"]["per_rank"]: print("  rank", r["rank"], r["ms_per_step_device"], r["exact_path_queries_per_step"])
    for s in j["config"]["secondary"]:
        print("   ", s.get("name","")[:40], s.get("seconds"), s.get("equals_reference_on_sample",{}).get("ok"), s.get("error"))
except Exception as ex:
    print("no N=4 result", ex)
PY
