timeout 600 python bench.py --batch 8 --erp 1024x2048 --nrows 5 --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/cfg3.json 2>gpurun_out/cfg3.err
timeout 600 python bench.py --batch 16 --erp 512x1024 --nrows 6 --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/cfg4.json 2>gpurun_out/cfg4.err
timeout 600 python bench.py --batch 1 --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/b1.json 2>gpurun_out/b1.err
