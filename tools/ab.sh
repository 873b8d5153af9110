timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "parity|passed|failed|Error|error" > gpurun_out/parity.log
for b in 8 32; do
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/new_${b}.json 2>>gpurun_out/ab.err
done
