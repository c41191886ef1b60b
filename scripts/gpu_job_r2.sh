set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -c 400 gpurun_out/bench_r2_n2.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench_r2_n2.json').read().strip().splitlines()[-1])
    print("N=2 value", round(j["value"]), "e2e", round(j["e2e"]["value"]), j["clocks"])
    for r in j["config"]["per_rank"]: print("  rank", r["rank"], r["ms_per_step_device"], r["exact_path_queries_per_step"])
    for s in j["config"]["secondary"]:
        print("   ", s.get("name","")[:40], s.get("seconds"), s.get("equals_reference_on_sample",{}).get("ok"), s.get("error"))
except Exception as ex:
    print("no N=2 result", ex)
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | cut -c1-300
