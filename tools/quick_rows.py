"""Per-row table of one library build, GPU side only.  usage: python tools/quick_rows.py <lib.so>"""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.abspath(sys.argv[1])
env = dict(os.environ, HEHUB_B200_LIB=lib)
out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu", "--sweep-cts", "0", "--steps", "10", "--warmup", "3"],
                     capture_output=True, text=True, env=env, cwd=root).stdout
d = json.loads(out[out.index("{"):])
rows = d["extras"]["rows_c3_shape"]["rows"]
print(os.path.basename(lib), " | ".join(f"{k.split()[0]} {k.split()[1][:18]} {v['frac_hbm']:.3f}" for k, v in rows.items() if k.startswith(("a5", "a6", "a7", "a8"))))
