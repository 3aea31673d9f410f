mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r1_gpu.txt 2>&1
timeout 60 tools/tmem_probe > gpurun_out/r1_tmem_probe.txt 2>&1
timeout 120 tools/mma_rate > gpurun_out/r1_mma_rate.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r1_pytest.log
timeout 600 python bench.py > gpurun_out/r1_bench_default.json 2> gpurun_out/r1_bench_default.err
timeout 600 python bench.py --batch 1024 --steps 300 > gpurun_out/r1_bench_b1024.json 2> gpurun_out/r1_bench_b1024.err
timeout 600 python bench.py --batch 4096 --steps 100 > gpurun_out/r1_bench_b4096.json 2> gpurun_out/r1_bench_b4096.err
tail -3 gpurun_out/r1_pytest.log
