mkdir -p gpurun_out/A
O=gpurun_out/A
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,driver_version --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_default_steps20.json 2> $O/bench_default_steps20.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 600 python bench.py --arch squeeze-redconv --precision fp16 --batch 1024 --steps 300 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 600 python bench.py --precision int8 --batch 4096 --steps 60 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
timeout 600 python bench.py --arch squeeze-redconv --precision int8 --batch 1024 --steps 200 > $O/bench_redconv_int8.json 2> $O/bench_redconv_int8.err
timeout 600 python bench.py --precision fp32 --steps 100 > $O/bench_fp32.json 2> $O/bench_fp32.err
ERNET_FUSE_INGEST=1 timeout 600 python bench.py > $O/bench_default_fused.json 2> $O/bench_default_fused.err
# ncu launch list of the bench command (per-launch durations; shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launch_list.csv python bench.py --steps 2 --warmup 3 > $O/ncu_launch_list.out 2>&1
# ncu --set full summaries per configuration (reports stay on the box; summaries come back)
mkdir -p /tmp/rep
timeout 900 ncu --set full --clock-control none -s 14 -c 7 -o /tmp/rep/bf16 python tools/run_forward.py squeeze-ernet bf16 256 4 > $O/ncu_bf16.out 2>&1
timeout 900 ncu --set full --clock-control none -s 24 -c 9 -o /tmp/rep/cfg3 python tools/run_forward.py squeeze-redconv fp16 1024 4 > $O/ncu_cfg3.out 2>&1
timeout 900 ncu --set full --clock-control none -s 40 -c 24 -o /tmp/rep/cfg4 python tools/run_forward.py squeeze-ernet int8 4096 3 > $O/ncu_cfg4.out 2>&1
timeout 900 ncu --set full --clock-control none -s 30 -c 12 -o /tmp/rep/fp32 python tools/run_forward.py squeeze-ernet fp32 256 4 > $O/ncu_fp32.out 2>&1
cp profiles/ncu_dram_bytes_per_launch.json /tmp/rep/old.json 2>/dev/null
python tools/ncu_summarize.py /tmp/rep/bf16.ncu-rep r02_step_bf16_b256 256 squeeze-ernet bf16 > /dev/null 2>&1
python tools/ncu_summarize.py /tmp/rep/cfg3.ncu-rep r02_step_redconv_fp16_b1024 1024 squeeze-redconv fp16 > /dev/null 2>&1
python tools/ncu_summarize.py /tmp/rep/cfg4.ncu-rep r02_step_int8_b1024 1024 squeeze-ernet int8 > /dev/null 2>&1
python tools/ncu_summarize.py /tmp/rep/fp32.ncu-rep r02_step_fp32_b256 256 squeeze-ernet fp32 > /dev/null 2>&1
cp profiles/r02_step_*_ncu_full.txt profiles/ncu_dram_bytes_per_launch.json $O/ 2>/dev/null
# race / sync / memory sanitizer on one small forward per engine
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_check.py > $O/sanitizer_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_check.py > $O/sanitizer_synccheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_check.py bf16 > $O/sanitizer_racecheck.log 2>&1
ls -la $O
