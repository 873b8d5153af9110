timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "parity\] (tensor|fused)|passed|failed|Error|error" | tail -8 > gpurun_out/tests.log
for b in 8 32; do
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/new_${b}.json 2>>gpurun_out/ab.err
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline --opt lanes=1 > gpurun_out/new_${b}_l1.json 2>>gpurun_out/ab.err
done
