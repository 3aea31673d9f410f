mkdir -p gpurun_out
ERNET_FUSE_INGEST=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graph or frames_path or fused or fast_ingest or host_submit or full_size" 2>&1 | tail -3 > gpurun_out/r7_pytest_fused.log
tail -2 gpurun_out/r7_pytest_fused.log
ERNET_FUSE_INGEST=1 timeout 300 python bench.py --steps 500 > gpurun_out/r7_bench_fused.json 2> gpurun_out/r7_bench_fused.err
ERNET_FUSE_INGEST=1 timeout 300 python bench.py --steps 200 --batch 1024 > gpurun_out/r7_bench_fused_b1024.json 2> gpurun_out/r7_bench_fused_b1024.err
ERNET_FUSE_INGEST=1 timeout 300 python bench.py --arch squeeze-redconv --precision fp16 --batch 1024 --steps 200 > gpurun_out/r7_bench_cfg3_fused.json 2> gpurun_out/r7_bench_cfg3_fused.err
ERNET_FUSE_INGEST=1 timeout 300 python bench.py --precision int8 --batch 4096 --steps 50 > gpurun_out/r7_bench_cfg4_fused.json 2> gpurun_out/r7_bench_cfg4_fused.err
