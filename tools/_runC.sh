mkdir -p gpurun_out/C /tmp/rep
O=gpurun_out/C
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "int8 or real_frames or fused_transform_block1 or graph" 2>&1 | grep -v "^$" | tail -30 > $O/pytest_int8.log
timeout 600 python -m pytest tests/test_next_rows.py -m gpu -q -s -k "build_trt or device_jpeg or evaluate" 2>&1 | tail -6 >> $O/pytest_int8.log
cat $O/pytest_int8.log | cut -c1-300
timeout 600 python bench.py --arch squeeze-redconv --precision int8 --batch 1024 --steps 200 > $O/bench_redconv_int8.json 2> $O/bench_redconv_int8.err
timeout 900 python tools/next_rows_time.py > $O/next_rows_time.jsonl 2> $O/next_rows_time.err
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/rep/b1024 python tools/run_forward.py squeeze-ernet bf16 1024 3 > $O/ncu_b1024.out 2>&1
python tools/ncu_summarize.py /tmp/rep/b1024.ncu-rep r02_step_bf16_b1024 1024 squeeze-ernet bf16 > /dev/null 2>> $O/ncu_b1024.out
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/rep/ri8 python tools/run_forward.py squeeze-redconv int8 1024 3 > $O/ncu_ri8.out 2>&1
python tools/ncu_summarize.py /tmp/rep/ri8.ncu-rep r02_step_redconv_int8_b1024 1024 squeeze-redconv int8 > /dev/null 2>> $O/ncu_ri8.out
cp profiles/r02_step_bf16_b1024_ncu_full.txt profiles/r02_step_redconv_int8_b1024_ncu_full.txt profiles/ncu_dram_bytes_per_launch.json $O/ 2>/dev/null
cat $O/next_rows_time.jsonl
