P='import json,sys; d=json.loads(sys.stdin.read()); print(TAG, round(d["value"]), {k: round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()})'
timeout 200 python -m pytest tests/test_pipeline_gpu.py -x -q 2>&1 | tail -3
for dbg in 0 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipe-chunk ${CHUNK:-2048} --pipe-debug $dbg 2>&1 | tail -1 | python -c "TAG='debug=$dbg'; $P"
done
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ivfadc_pipe -s 8 -c 1 -o gpurun_out/$NCU python bench.py --steps 1 --warmup 3 --no-cpu-baseline --pipe-chunk 2048 > gpurun_out/ncu_pipe.log 2>&1
tail -2 gpurun_out/ncu_pipe.log
fi
