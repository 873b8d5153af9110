timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests.log
for b in 8 32; do
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/new_${b}.json 2>>gpurun_out/ab.err
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline --opt cta2=0 > gpurun_out/new_${b}_cta20.json 2>>gpurun_out/ab.err
done
timeout 300 python tools/probe_tail.py > gpurun_out/probe_tail.log 2>&1
