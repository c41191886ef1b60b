set -x
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/gputests.log 2>&1; tail -5 gpurun_out/gputests.log
for L in 0 8; do PF_LOCKSTEP=$L timeout 200 ncu --set full --clock-control none --import-source on -k regex:prefilter_gemm -s 1 -c 1 -o gpurun_out/r2_prof_prefilter_L$L scripts/umma_probe 1000 3000000 300 1 perf > gpurun_out/ncu_prefilter_L$L.log 2>&1; tail -2 gpurun_out/ncu_prefilter_L$L.log; done
timeout 900 python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 400 gpurun_out/bench_r2c.err
