for b in 8; do
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/new_${b}.json 2>>gpurun_out/ab.err
done
timeout 300 ncu --set full --clock-control none -k regex:heads_kernel -c 1 -o gpurun_out/heads python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
