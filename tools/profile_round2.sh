set -x
python bench.py --steps 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_n1_reference_arm.json 2> gpurun_out/ref.err
python bench.py --steps 5 --engine lane-t --no-cpu-baseline > gpurun_out/r2_bench_n1_lane_t.json 2>/dev/null
python bench.py --workload cfg3 --steps 5 > gpurun_out/r2_bench_cfg3.json 2>/dev/null
python bench.py --workload cfg4 --steps 3 > gpurun_out/r2_bench_cfg4.json 2>/dev/null
python bench.py --workload cfg5 --steps 2 --batch 32 > gpurun_out/r2_bench_cfg5_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_full.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lane -s 6 -c 2 -o gpurun_out/r2_lane_staged python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lane -s 6 -c 2 -o gpurun_out/r2_lane_t python bench.py --steps 1 --warmup 1 --no-cpu-baseline --engine lane-t > /dev/null 2>&1
ls -la gpurun_out/
tail -c 600 gpurun_out/r2_bench_n1.json
tail -c 400 gpurun_out/r2_bench_n1_reference_arm.json
