mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_pytest.log
timeout 300 python bench.py --steps 500 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
ERNET_PF_EPI=0 timeout 300 python bench.py --steps 500 > gpurun_out/r2_bench_default_oldepi.json 2> gpurun_out/r2_bench_default_oldepi.err
timeout 300 python bench.py --arch squeeze-redconv --precision fp16 --batch 1024 --steps 200 > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err
ERNET_PF_EPI=0 timeout 300 python bench.py --arch squeeze-redconv --precision fp16 --batch 1024 --steps 200 > gpurun_out/r2_bench_cfg3_oldepi.json 2> gpurun_out/r2_bench_cfg3_oldepi.err
timeout 300 python bench.py --precision int8 --batch 4096 --steps 50 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err
tail -3 gpurun_out/r2_pytest.log
