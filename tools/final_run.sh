set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3_tests.log 2>&1; tail -3 gpurun_out/r3_tests.log
timeout 900 python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; tail -c 600 gpurun_out/r3_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3_ref.json 2>> gpurun_out/r3_bench.err
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r3_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swin_attn2 -s 5 -c 1 -f -o gpurun_out/r3_attn python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r3_ncu2.log 2>&1
ls -la gpurun_out/r3_*
