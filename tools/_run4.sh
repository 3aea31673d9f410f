mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s -k "not (int8 and redconv)" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r4_pytest.log
tail -4 gpurun_out/r4_pytest.log
timeout 300 python bench.py --steps 500 > gpurun_out/r4_bench_default.json 2> gpurun_out/r4_bench_default.err
ERNET_FUSE_INGEST=0 timeout 300 python bench.py --steps 500 > gpurun_out/r4_bench_default_unfused.json 2> gpurun_out/r4_bench_default_unfused.err
timeout 300 python bench.py --arch squeeze-redconv --precision fp16 --batch 1024 --steps 200 > gpurun_out/r4_bench_cfg3.json 2> gpurun_out/r4_bench_cfg3.err
timeout 300 python bench.py --precision int8 --batch 4096 --steps 50 > gpurun_out/r4_bench_cfg4.json 2> gpurun_out/r4_bench_cfg4.err
