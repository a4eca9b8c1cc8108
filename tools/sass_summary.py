"""Per-kernel SASS mnemonic summary of libnnic.so (cuobjdump -sass): the Blackwell-native instructions each kernel contains
(UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, SYNCS = mbarrier, ATOMS = shared
atomic, STG.E.ENL2.256 = 256-bit global store ...).  python tools/sass_summary.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neural_network_image_compression_b200", "libnnic.so")
WANT = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "ATOMS", "ATOMG", "REDG", "RED",
        "STG.E.ENL2.256", "LDG.E.ENL2.256", "STG.E.128", "LDG.E.128", "STS.128", "LDS.128", "FFMA", "HFMA2", "MUFU", "BAR.SYNC", "ELECT", "FENCE")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WANT:
                if op == w or op.startswith(w + ".") or (("." in w) and op.startswith(w)):
                    kernels[cur][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.basename(LIB)}: SASS mnemonics per kernel (cuobjdump -sass, sm_100a); counts are static instructions")
    for (name, cnt), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(anonymous namespace\)::|nnic::|void ", "", dn).split("(")[0]
        items = " ".join(f"{k}={v}" for k, v in cnt.items() if k != "_total" and v)
        print(f"{short:60s} instr={cnt['_total']:5d}  {items}")


if __name__ == "__main__":
    main()
