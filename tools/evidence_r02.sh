set -x
cd gpurun_out
timeout 600 python -m pytest ../tests -m gpu -q -s > r2_gpu_tests_v36.log 2>&1; tail -2 r2_gpu_tests_v36.log
cd ..
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v36.json 2> gpurun_out/bench.err
python bench.py --steps 20 --warmup 5 --batch 32 --skip-cpu-baseline > gpurun_out/r2_bench_v36_b32.json 2>> gpurun_out/bench.err
python bench.py --steps 50 --warmup 5 --batch 1 --skip-cpu-baseline > gpurun_out/r2_bench_v36_b1.json 2>> gpurun_out/bench.err
python bench.py --steps 20 --warmup 5 --batch 8 --erp 1024x2048 --nrows 5 --skip-cpu-baseline > gpurun_out/r2_bench_v36_cfg3_1gpu.json 2>> gpurun_out/bench.err
python bench.py --steps 20 --warmup 5 --batch 16 --nrows 6 --skip-cpu-baseline > gpurun_out/r2_bench_v36_cfg4_1gpu.json 2>> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_v36_reference_arm.json 2>> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python tools/timeline.py > gpurun_out/r2_timeline_b8_graph_v36.log 2>&1; tail -2 gpurun_out/r2_timeline_b8_graph_v36.log
python tools/probe_token.py > gpurun_out/r2_token_stack_probe.log 2>&1
python tools/probe_rolling.py > gpurun_out/r2_probe_rolling_nstack.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/r2_launches_v36.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/r2_launch_bench.log 2>&1
du -sh gpurun_out
