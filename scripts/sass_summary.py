"""Per-kernel SASS instruction summary of libillico_b200.so (evidence for profiles/: which kernels use the TMA engine,
mbarriers, shared-memory atomics, FP64 ...).

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "illico_b200", "libillico_b200.so")
KEYS = ["UBLKCP", "SYNCS", "LDS", "STS", "ATOMS", "ATOMG", "RED", "LDG", "STG", "LDL", "STL", "DFMA", "DMUL", "DADD", "MUFU",
        "IMAD", "SHFL", "VOTE", "MATCH", "BAR", "FADD", "ISETP", "FSETP", "BRA"]


def demangle(name):
    try:
        return subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in ln:
            usage[cur] = " ".join(re.findall(r"(?:REG|STACK|SHARED|LOCAL):\d+", ln))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            kernels[cur][op.split(".")[0]] += 1
            if op.startswith("UBLKCP") or op.startswith("SYNCS") or op.startswith("ATOMS") or op.startswith("LDS.U16") or op.startswith("STS.U16"):
                kernels[cur][op] += 1
    print("kernel | resources | SASS instructions | selected mnemonics (static counts)")
    for k, c in kernels.items():
        name = demangle(k)
        name = re.sub(r"illico::\(anonymous namespace\)::|illico::", "", name)
        name = re.sub(r"\((?:int|bool)\)", "", name).replace("<unnamed>::", "")
        name = name.split("(")[0].replace("void ", "")
        sel = ", ".join(f"{m} {c[m]}" for m in KEYS if c.get(m))
        extra = ", ".join(f"{m} {n}" for m, n in c.items() if "." in m)
        print(f"{name} | {usage.get(k, '')} | {c['_total']} | {sel}" + (f" | {extra}" if extra else ""))


if __name__ == "__main__":
    sys.exit(main())
