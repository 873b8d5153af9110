timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "parity|passed|failed|Error" > gpurun_out/parity.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_final.err
timeout 300 python bench.py --batch 32 --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_b32.json 2>> gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/b_ncu.log 2>&1
i=0
for sk in "9 2" "62 10"; do set -- $sk; i=$((i+1))
timeout 400 ncu --set full --clock-control none -k regex:conv_tc --launch-skip $1 -c $2 -o gpurun_out/conv_full_$i python tools/one_forward.py 8 1 > gpurun_out/ncu_full_$i.log 2>&1
ncu -i gpurun_out/conv_full_$i.ncu-rep --page raw --csv > gpurun_out/conv_full_$i.csv 2>/dev/null
rm -f gpurun_out/conv_full_$i.ncu-rep
done
du -sh gpurun_out
