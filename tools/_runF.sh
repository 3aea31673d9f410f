mkdir -p gpurun_out/F
O=gpurun_out/F
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n${N}_steps20.json 2> $O/bench_n${N}_steps20.err
tail -c 600 $O/bench_n${N}_steps20.err
python tools/summarize_bench.py $O/bench_ref_n$N.json $O/bench_n${N}_steps20.json
