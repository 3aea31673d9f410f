mkdir -p gpurun_out/B2
O=gpurun_out/B2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_steps20.json 2> $O/bench_n2_steps20.err
timeout 600 $TR bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 600 $TR bench.py --gpus 2 --batch 256 > $O/bench_n2_weak256.json 2> $O/bench_n2_weak256.err
timeout 600 python tools/eval_sharded_check.py > $O/eval_sharded.log 2>&1
tail -2 $O/*.err
ls -la $O
