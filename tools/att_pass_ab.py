"""A/B of the attention head passes (OQ_ATT_PASS_MB) on the LLaMA-7B step: graph-replayed step time per setting."""
import os
import subprocess
import sys

for mb in sys.argv[1:] or ["0", "72", "40", "140"]:
    env = dict(os.environ, OQ_ATT_PASS_MB=mb)
    out = subprocess.run([sys.executable, "tools/step_time.py"], env=env, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("graph train_step")]
    print(f"OQ_ATT_PASS_MB={mb}: {line[0] if line else out[-300:]}", flush=True)
