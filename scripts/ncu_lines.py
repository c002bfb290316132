"""Aggregates an ncu SASS source page by CUDA source line (ncu's CLI prints per-line metrics only for SASS).

    python scripts/ncu_lines.py <report.ncu-rep> <kernel substring> <cubin from cuobjdump -xelf> [top N]

Line info comes from `nvdisasm -g -c` of the cubin (built with -lineinfo); SASS rows are matched in order.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kname, cubin = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kname}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    body = []
    for r in rows[start + 1:]:
        if r and r[0] == "Kernel Name":
            break  # only the first matching launch
        if r and r[0].startswith("0x"):
            body.append(r)
    ci = {h: i for i, h in enumerate(hdr)}
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # collect (line, opcode text) for every instruction of sections whose name contains kname (and device functions)
    ins = []
    cur_line, cur_file, in_text = None, None, False
    for ln in dis:
        if ln.startswith(".text."):
            in_text = True
        elif ln.startswith(".section") or ln.startswith(".nv"):
            pass
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_file, cur_line = m.group(1).split("/")[-1], int(m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and in_text:
            ins.append((cur_file, cur_line, m.group(2).strip()))
    # align: find the offset in `ins` where the first few opcodes of the ncu listing match
    def op(s):
        return s.strip().split()[0] if s.strip() else ""
    first = [op(r[ci["Source"]]) for r in body[:12]]
    off = None
    for k in range(len(ins) - len(body) + 1):
        if [op(x[2]) for x in ins[k:k + 12]] == first:
            off = k
            break
    if off is None:
        print("could not align SASS with nvdisasm output; printing SASS hot spots only")
    agg = defaultdict(lambda: [0, 0, 0])
    tot_inst = tot_samp = 0
    for i, r in enumerate(body):
        inst = int(r[ci["Instructions Executed"]] or 0)
        samp = int(r[ci["# Samples"]] or 0)
        key = (ins[off + i][0], ins[off + i][1]) if off is not None and off + i < len(ins) else ("?", i)
        agg[key][0] += inst
        agg[key][1] += samp
        agg[key][2] += int(r[ci["Thread Instructions Executed"]] or 0)
        tot_inst += inst
        tot_samp += samp
    print(f"kernel {kname}: {len(body)} SASS instructions, {tot_inst} warp instructions executed, {tot_samp} samples")
    print(f"{'file:line':28s} {'warp inst':>12s} {'%':>6s} {'samples':>9s} {'%':>6s} {'thr/inst':>8s}")
    for key, (inst, samp, thr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{str(key[0]) + ':' + str(key[1]):28s} {inst:12d} {100 * inst / max(tot_inst, 1):6.2f} {samp:9d} "
              f"{100 * samp / max(tot_samp, 1):6.2f} {thr / max(inst, 1):8.1f}")


if __name__ == "__main__":
    main()
