import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for which, pads in (("lte", (0, 60000, 100000)), ("n64", (0, 30000, 60000, 120000))):
    for pad in pads:
        env = dict(os.environ, OFDM_SMEM_PAD=str(pad))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_bench.py"), which], env=env, capture_output=True, text=True).stdout
        for l in out.splitlines():
            try: d = json.loads(l)
            except Exception: continue
            print("pad %6d  %-10s rx %.3f ms  tx %.3f ms" % (pad, d["name"], d["rx_ms"], d["tx_ms"]))
