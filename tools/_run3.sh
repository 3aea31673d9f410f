mkdir -p gpurun_out
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
ncu --set full --import-source on --clock-control none -k regex:acff_pblock -s 2 -c 1 -o gpurun_out/r3_pf python /tmp/one.py > gpurun_out/r3_ncu_pf.log 2>&1
ERNET_PF_EPI=0 ncu --set full --import-source on --clock-control none -k regex:acff_pblock -s 2 -c 1 -o gpurun_out/r3_old python /tmp/one.py > gpurun_out/r3_ncu_old.log 2>&1
ls -la gpurun_out/*.ncu-rep
