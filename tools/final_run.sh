# Round-end evidence on one B200: full GPU test suite, smoke(), the bench lines of profiles/r02_bench_*.json(l).
set -x
mkdir -p gpurun_out/final
nvidia-smi --query-gpu=name,clocks.max.sm,driver_version --format=csv > gpurun_out/final/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/final/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err
python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_default_steps20.json 2> gpurun_out/final/bench_default_steps20.err
if [ "$1" = "all" ]; then
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final/bench_reference_arm.json 2> gpurun_out/final/bench_reference_arm.err
: > gpurun_out/final/bench_other_configs.jsonl
python bench.py --arch squeeze-redconv --precision fp16 --batch 1024 --steps 300 --warmup 20 >> gpurun_out/final/bench_other_configs.jsonl 2> gpurun_out/final/o1.err
python bench.py --arch squeeze-ernet --precision int8 --batch 4096 --steps 60 --warmup 20 >> gpurun_out/final/bench_other_configs.jsonl 2> gpurun_out/final/o2.err
python bench.py --arch squeeze-redconv --precision int8 --batch 1024 --steps 200 --warmup 20 >> gpurun_out/final/bench_other_configs.jsonl 2> gpurun_out/final/o3.err
python bench.py --precision fp32 --steps 100 --warmup 20 >> gpurun_out/final/bench_other_configs.jsonl 2> gpurun_out/final/o4.err
ERNET_FUSE_INGEST=1 python bench.py --steps 2000 --warmup 20 >> gpurun_out/final/bench_other_configs.jsonl 2> gpurun_out/final/o5.err
ERNET_TAIL_TILES=0 python bench.py --steps 2000 --warmup 20 >> gpurun_out/final/bench_other_configs.jsonl 2> gpurun_out/final/o6.err
fi
python - <<'PY'
import json, os
for f in ("bench_default.json","bench_default_steps20.json","bench_reference_arm.json"):
    try:
        d=json.load(open("gpurun_out/final/"+f)); print(f, round(d["value"]), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
if os.path.exists("gpurun_out/final/bench_other_configs.jsonl"):
    for l in open("gpurun_out/final/bench_other_configs.jsonl"):
        d=json.loads(l); print(d["config"]["workload"][:60], round(d["value"]), d["ms_per_step"])
PY
