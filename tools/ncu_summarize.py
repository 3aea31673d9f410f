"""Turn an ncu report (--set full) of tools/run_forward.py into the committed summaries under profiles/.

    python tools/ncu_summarize.py <report.ncu-rep> <tag> <batch> [arch] [precision]

Writes profiles/<tag>_ncu_full.txt and merges the per-launch DRAM bytes / tensor-pipe activity of the configuration
into profiles/ncu_dram_bytes_per_launch.json under configs["<arch>/<precision>/<batch>"] (bench.py's roofline.traffic)."""
import csv
import json
import subprocess
import sys

rep, tag, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
arch = sys.argv[4] if len(sys.argv) > 4 else "squeeze-ernet"
precision = sys.argv[5] if len(sys.argv) > 5 else "bf16"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem"]
stage_of = [("ingest_block1", "tc_block1"), ("ingest_stem", "ingest"), ("ingest_kernel", "ingest"), ("PCfg<12, 64", "red2"),
            ("PCfg<2, 64", "tc_block1"), ("CCfg<2, 64", "tc_block1"), ("BlockCfg<2, 64", "tc_block1"),
            ("CCfg<8, 96", "tc_block2"), ("CCfg<4, 96", "tc_block2"), ("PCfg<8, 96", "tc_block2"), ("BlockCfg<8, 96", "tc_block2"), ("BlockCfg<4, 96", "tc_block2"),
            ("CCfg<12, 128", "tc_block3"), ("CCfg<6, 128", "tc_block3"), ("CCfg<4, 128", "tc_block3"), ("PCfg<12, 128", "tc_block3"),
            ("BlockCfg<12, 128", "tc_block3"), ("BlockCfg<6, 128", "tc_block3"), ("acff4_head", "tc_block4"),
            ("acff_dw", "dw#"), ("pointwise_kernel", "pw#"), ("pw32_kernel", "pw#"), ("stem_kernel", "stem"), ("head_kernel", "head")]
seen = {}


def stage_name(name, red):
    st = next((v for p, v in stage_of if p in name), None)
    if st and st.endswith("#"):                       # repeated layer-wise kernels: number them in launch order
        k = st[:-1]
        seen[k] = seen.get(k, 0) + 1
        if k == "pw" and red:                         # Squeeze_RedConv: pw1, pw2, red2, pw3, red3, pw4
            return ["pw1", "pw2", "red2", "pw3", "red3", "pw4"][min(seen[k], 6) - 1]
        return f"{k}{seen[k]}"
    return st


lines = [f"# ncu --set full --clock-control none, tools/run_forward.py, batch {batch} ({tag})\n",
         "# per-launch values from cold-cache, serialised replays: compare shares, not absolutes\n"]
dram = {}
tpipe = {}
tot = 0.0
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    lines.append("\n## " + name.split("(")[0][:110] + "\n")
    for k in keys:
        if k in hdr:
            lines.append(f"{k} = {r[hdr.index(k)]} {units[hdr.index(k)]}\n")
    st = stage_name(name, arch == "squeeze-redconv")

    def to_bytes(k):
        v, u = float(r[hdr.index(k)]), units[hdr.index(k)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    if st and st not in dram:
        dram[st] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        tpipe[st] = float(r[hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")])
    tot += float(r[hdr.index("gpu__time_duration.sum")])
lines.append(f"\n# sum of kernel durations in this capture: {tot:.1f} us\n")
open(f"profiles/{tag}_ncu_full.txt", "w").writelines(lines)
import os
path = "profiles/ncu_dram_bytes_per_launch.json"
doc = json.load(open(path)) if os.path.exists(path) else {}
if "configs" not in doc:
    doc = {"configs": {}}
doc["configs"][f"{arch}/{precision}/{batch}"] = {"batch": batch, "source": f"profiles/{tag}_ncu_full.txt", "stages": dram, "tensor_pipe_active_pct": tpipe}
json.dump(doc, open(path, "w"), indent=1)
print("".join(lines[-40:]))
