#!/usr/bin/env python
"""One line per bench JSON file: value, ms/step, e2e (and its share of the H2D ceiling), parity, stage times, roofline."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:          # noqa: BLE001
        print(f"{path}: unreadable ({e})")
        continue
    if d.get("impl") == "reference":
        print(f"{path}: reference arm {d['value']:.1f} img/s, {d['ms_per_step']:.0f} ms/step, steps {d['steps']}, cores {d['cpu_baseline']['cores']}")
        continue
    r, e = d["roofline"], d["e2e"]
    print(f"{path}: N={d['n_gpus']} {d['dtype']} value {d['value']:.0f} img/s ({d['ms_per_step']:.4f} ms/step, steps {d['steps']}), "
          f"e2e {e['value']:.0f} ({e.get('frac_of_h2d_ceiling', 0):.3f} of the H2D ceiling {e.get('h2d_ceiling_gbs', 0):.1f} GB/s), "
          f"parity err {d['parity']['rel_err']:.2e} flips {d['parity']['flips']}/{d['parity']['n']}, clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
    print(f"    dominant {r['kernel']}: {r['launch_ms'] * 1e3:.1f} us, {r['achieved']:.1f} {r['unit']} = {r['frac']:.3f} of {r['peak']:.0f}"
          f"{' (issued %.3f)' % r['issued_frac'] if 'issued_frac' in r else ''}, traffic {r['traffic']}, stages {r['stage_ms_per_step']}")
    if d.get("gather_probabilities_ms") is not None:
        print(f"    gather of the probabilities: {d['gather_probabilities_ms'] * 1e3:.1f} us")
    print(f"    depthwise_hbm block1 {d['depthwise_hbm']['block1']['frac']} block2 {d['depthwise_hbm']['block2']['frac']}; latency b1 {d['latency_b1_ms'] * 1e3:.1f} us, graph {d['latency_b1_graph_ms'] * 1e3:.1f} us; "
          + (f"cpu_baseline {d['cpu_baseline']['value']:.0f} img/s on {d['cpu_baseline']['cores']} cores" if d.get("cpu_baseline") else "cpu_baseline: N=1 only"))
