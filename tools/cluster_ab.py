import json, subprocess, sys, os
for rep in range(2):
    for c in ("0", "1", "2"):
        env = dict(os.environ, NNIC_TC_CLUSTER=c)
        out = subprocess.run([sys.executable, "bench.py", "--steps", "60", "--no-cpu-baseline"], capture_output=True, text=True, env=env).stdout
        d = json.loads(out)
        k = d["kernels"]
        print("cluster=%s value %.0f ms %.4f  conv4 %.4f dconv6 %.4f conv3 %.4f dconv5 %.4f dconv7 %.4f conv2 %.4f" % (c, d["value"], d["ms_per_step"], k["conv4"]["ms_per_launch"], k["dconv6"]["ms_per_launch"], k["conv3"]["ms_per_launch"], k["dconv5"]["ms_per_launch"], k["dconv7"]["ms_per_launch"], k["conv2"]["ms_per_launch"]), flush=True)
