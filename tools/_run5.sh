mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "int8 or real or graph or frames_path or fused or fast_ingest or host_submit or tc_engine or schedules" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r5_pytest.log
tail -4 gpurun_out/r5_pytest.log
timeout 300 python bench.py --steps 500 > gpurun_out/r5_bench_default.json 2> gpurun_out/r5_bench_default.err
timeout 300 python bench.py --arch squeeze-redconv --precision int8 --batch 1024 --steps 100 > gpurun_out/r5_bench_red_int8.json 2> gpurun_out/r5_bench_red_int8.err
cat > /tmp/one.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import fixtures, rtdm_b200
dev = torch.device('cuda:0')
m = rtdm_b200.from_state_dict('squeeze-ernet', fixtures.get_state_dict('squeeze-ernet', 'shipped'), dev, 'bf16')
f = torch.randint(0, 256, (256, 240, 240, 3), dtype=torch.uint8, device=dev)
for _ in range(3):
    m.forward_frames(f)
torch.cuda.synchronize()
PY
ncu --set full --import-source on --clock-control none -k regex:ingest_block1 -s 2 -c 1 -o gpurun_out/r5_fused python /tmp/one.py > gpurun_out/r5_ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep
