timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/tests.log
for b in 8 32; do
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/new_${b}.json 2>>gpurun_out/ab.err
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline --opt direct32=1 > gpurun_out/new_${b}_d32.json 2>>gpurun_out/ab.err
done
timeout 300 python tools/probe_tail.py 2>&1 | tail -75 > gpurun_out/probe_tail.log
