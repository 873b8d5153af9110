set -x
cd gpurun_out
timeout 600 python -m pytest ../tests -m gpu -q -s > r2_gpu_tests_v33.log 2>&1; tail -2 r2_gpu_tests_v33.log
cd ..
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v33.json 2> gpurun_out/bench.err
python bench.py --steps 20 --warmup 5 --batch 32 --skip-cpu-baseline > gpurun_out/r2_bench_v33_b32.json 2>> gpurun_out/bench.err
python bench.py --steps 50 --warmup 5 --batch 1 --skip-cpu-baseline > gpurun_out/r2_bench_v33_b1.json 2>> gpurun_out/bench.err
python bench.py --steps 20 --warmup 5 --batch 8 --erp 1024x2048 --nrows 5 --skip-cpu-baseline > gpurun_out/r2_bench_v33_cfg3_1gpu.json 2>> gpurun_out/bench.err
python bench.py --steps 20 --warmup 5 --batch 16 --nrows 6 --skip-cpu-baseline > gpurun_out/r2_bench_v33_cfg4_1gpu.json 2>> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python tools/timeline.py > gpurun_out/r2_timeline_b8_graph_v33.log 2>&1; tail -2 gpurun_out/r2_timeline_b8_graph_v33.log
python tools/probe_token.py > gpurun_out/r2_token_stack_probe.log 2>&1
python tools/probe_rolling.py > gpurun_out/r2_probe_rolling_nstack.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/r2_launches_v33.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/r2_launch_bench.log 2>&1
cd gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"token_stack|conv_tc_kernel<16" -c 4 -o r2_tok_v33 python ../tools/one_forward.py 8 1 > r2_tok_v33.log 2>&1; ncu -i r2_tok_v33.ncu-rep --page raw --csv > r2_ncu_token_v33_raw.csv; rm -f r2_tok_v33.ncu-rep
cd ..
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_token_stack.py -q -x -k "whole_stack and (18-8 or 46-3 or 18-23) or phase_by_phase" > gpurun_out/r2_memcheck_token_stack.log 2>&1; tail -4 gpurun_out/r2_memcheck_token_stack.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_token_stack.py -q -x -k "whole_stack and 18-1" > gpurun_out/r2_racecheck_token_stack.log 2>&1; tail -4 gpurun_out/r2_racecheck_token_stack.log
du -sh gpurun_out
