mkdir -p gpurun_out
ERNET_FUSE_INGEST=1 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "graph or frames_path or fused or fast_ingest or host_submit or full_size or config3" 2>&1 | grep -v "^$" | tail -15 > gpurun_out/r6_pytest_fused.log
tail -3 gpurun_out/r6_pytest_fused.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graph or frames_path or fused or fast_ingest" 2>&1 | tail -3 > gpurun_out/r6_pytest_unfused.log
tail -2 gpurun_out/r6_pytest_unfused.log
timeout 300 python bench.py --steps 500 > gpurun_out/r6_bench_unfused.json 2> gpurun_out/r6_bench_unfused.err
ERNET_FUSE_INGEST=1 timeout 300 python bench.py --steps 500 > gpurun_out/r6_bench_fused.json 2> gpurun_out/r6_bench_fused.err
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
ERNET_FUSE_INGEST=1 ncu --set full --import-source on --clock-control none -k regex:ingest_block1 -s 2 -c 1 -o gpurun_out/r6_fused python /tmp/one.py > gpurun_out/r6_ncu_fused.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:ingest_stem5 -s 2 -c 1 -o gpurun_out/r6_ingest python /tmp/one.py > gpurun_out/r6_ncu_ingest.log 2>&1
