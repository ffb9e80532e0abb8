"""Render the per-row table of a bench.py JSON line as markdown.  usage: python tools/rows_table.py bench.json > table.md"""
import json, sys
txt = open(sys.argv[1]).read()
d = json.loads(txt[txt.index("{"):])
t = d["extras"]["rows_c3_shape"]
print(f"# SURVEY §8(a) rows, GPU and CPU side by side — shape N={t['shape']['N']}, L={t['shape']['L']} ({t['shape']['moduli_bits']} + P{t['shape']['special_bits']})\n")
print(t.get("note", ""), "\n")
print("| row | unit | GPU units/s | algorithmic GB/s | frac of HBM peak | reference CPU units/s (1 core) | GPU / 1 core |")
print("|---|---|---|---|---|---|---|")
for name, r in t["rows"].items():
    cpu = f"{r['cpu_per_s']:.4g}" if "cpu_per_s" in r else "–"
    sp = f"{r['gpu_over_one_core']:.0f}×" if "gpu_over_one_core" in r else "–"
    print(f"| {name} | {r['unit']} | {r['gpu_per_s']:.4g} | {r['gbs_algorithmic']:.0f} | {r['frac_hbm']:.3f} | {cpu} | {sp} |")
