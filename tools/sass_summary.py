"""SASS instruction-count summary of libpio_b200.so (evidence that the kernels are Blackwell-native):

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

Per kernel: tcgen05 MMAs (UTCHMMA, .2CTA = cta_group::2), TMEM loads / stores (LDTM / STTM), TMA tensor loads / stores
(UTMALDG / UTMASTG, multicast counted separately), bulk copies (UBLKCP), legacy warp MMAs (HMMA — must be zero),
MUFU.EX2 and the total instruction count.  Runs in the GPU-less build container (cuobjdump reads the embedded cubin)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "perceiverio_pytorch_b200", "libpio_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
            ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.MULTICAST", r"\bUTMALDG\S*MULTICAST"), ("UTMASTG", r"\bUTMASTG"),
            ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"),
            ("MUFU.EX2", r"\bMUFU\.EX2"), ("FFMA2", r"\bFFMA2"), ("global ATOM/RED", r"^(@!?U?P\d+\s+)?(ATOMG|ATOM|RED|REDG)\b")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        ins = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if not ins:
            continue
        text = ins.group(1)
        kernels[cur]["total"] += 1
        for name, rx in PATTERNS:
            if re.search(rx, text):
                kernels[cur][name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    cols = ["total"] + [n for n, _ in PATTERNS]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (sm_100a)")
    print("kernel | " + " | ".join(cols))
    tot = collections.Counter()
    for (k, c), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("pio::", "")
        print(short + " | " + " | ".join(str(c[n]) for n in cols))
        tot.update(c)
    print("ALL | " + " | ".join(str(tot[n]) for n in cols))
    if tot["HMMA"]:
        print("WARNING: legacy HMMA present", file=sys.stderr)


if __name__ == "__main__":
    main()
