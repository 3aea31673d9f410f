mkdir -p gpurun_out/B /tmp/rep
O=gpurun_out/B
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "real_frames or fused_transform_block1" 2>&1 | tail -5 > $O/pytest_subset.log
timeout 600 python -m pytest tests/test_next_rows.py -m gpu -q -k "build_trt" 2>&1 | tail -3 >> $O/pytest_subset.log
cat $O/pytest_subset.log
for cfg in "squeeze-ernet bf16 256 r02_step_bf16_b256" "squeeze-redconv fp16 1024 r02_step_redconv_fp16_b1024" "squeeze-ernet int8 1024 r02_step_int8_b1024" "squeeze-redconv int8 1024 r02_step_redconv_int8_b1024" "squeeze-ernet fp32 256 r02_step_fp32_b256"; do
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/rep/$4 python tools/run_forward.py $1 $2 $3 3 > $O/ncu_$4.out 2>&1
  python tools/ncu_summarize.py /tmp/rep/$4.ncu-rep $4 $3 $1 $2 > /dev/null 2>> $O/ncu_$4.out
done
cp profiles/r02_step_*_ncu_full.txt profiles/ncu_dram_bytes_per_launch.json $O/ 2>/dev/null
ERNET_FUSE_INGEST=1 timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/rep/fused python tools/run_forward.py squeeze-ernet bf16 256 3 > $O/ncu_fused.out 2>&1
python tools/ncu_summarize.py /tmp/rep/fused.ncu-rep r02_step_bf16_b256_fused 256 squeeze-ernet bf16-fused > /dev/null 2>> $O/ncu_fused.out
cp profiles/r02_step_bf16_b256_fused_ncu_full.txt $O/ 2>/dev/null
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
ls $O
