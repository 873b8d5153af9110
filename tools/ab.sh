timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/tests.log
for b in 8 32; do
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/new_${b}.json 2>>gpurun_out/ab.err
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline --opt khr_row64=0 > gpurun_out/new_${b}_r128.json 2>>gpurun_out/ab.err
done
