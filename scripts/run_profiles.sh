# round-1 (second session) evidence run: bench line, reference arm, launch list, full ncu of the dominant kernels
set -x
timeout 300 python bench.py > gpurun_out/r1s2_bench.json 2> gpurun_out/r1s2_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1s2_bench_reference.json 2>> gpurun_out/r1s2_bench.err
timeout 300 python bench.py --pipeline 0 --no-cpu-baseline > gpurun_out/r1s2_bench_separate_kernels.json 2>> gpurun_out/r1s2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1s2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ivfadc_pipe|coarse_select' -s 15 -c 5 -o gpurun_out/r1s2_pipe python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1s2_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'adc_scan_query|lut_build' -s 12 -c 2 -o gpurun_out/r1s2_separate python bench.py --steps 1 --warmup 3 --no-cpu-baseline --pipeline 0 >> gpurun_out/r1s2_ncu.log 2>&1
tail -3 gpurun_out/r1s2_ncu.log
cat gpurun_out/r1s2_bench.json | cut -c1-300
