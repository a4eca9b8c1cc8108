"""Per-kernel A/B of builds of libnnic on one box (interleaved rounds): python tools/lib_ab.py libA.so libB.so [rounds] [kernels...]"""
import json, os, subprocess, sys
libs, rounds = sys.argv[1:3], int(sys.argv[3]) if len(sys.argv) > 3 else 2
names = sys.argv[4:] or ["conv1", "dconv8", "conv2", "conv3", "conv4", "conv8", "dconv1", "dconv5", "dconv6", "dconv7"]
for rep in range(rounds):
    for lib in libs:
        env = dict(os.environ, NNIC_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, "bench.py", "--steps", "60", "--no-cpu-baseline", "--no-strong-c5"], capture_output=True, text=True, env=env).stdout
        d = json.loads(out)
        k = d["kernels"]
        print(os.path.basename(lib).ljust(18), "step %.4f ms |" % d["ms_per_step"], " ".join("%s %.4f" % (n, k[n]["ms_per_launch"]) for n in names), flush=True)
