timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q -s -k "heads_on_tensor" 2>&1 | grep -E "parity|passed|failed|Error|error|assert" | head -20 > gpurun_out/heads_tc.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/tests.log
for b in 8 32; do
timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/new_${b}.json 2>>gpurun_out/ab.err
done
