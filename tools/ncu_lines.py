"""Join an .ncu-rep SASS source page with nvdisasm line info: executed warp instructions and stall
samples per CUDA source line. Usage: ncu_lines.py report.ncu-rep object.o mangled_kernel_substring [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter


def main():
    rep, obj, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines_of_inst = []
    cur_line, active = None, False
    src_file = None
    for l in dis.splitlines():
        if l.startswith(".text."):
            active = kern in l
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur_line = int(m.group(2))
            src_file = m.group(1)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines_of_inst.append(cur_line)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[1]
    data = [r for r in rows[2:] if len(r) > h.index('Instructions Executed')]
    assert abs(len(data) - len(lines_of_inst)) <= 16, (len(data), len(lines_of_inst))  # trailing padding
    iN, iP = h.index('Instructions Executed'), h.index('# Samples')
    inst, samp = Counter(), Counter()
    for r, ln in zip(data, lines_of_inst):
        inst[ln] += int(r[iN] or 0)
        samp[ln] += int(r[iP] or 0)
    tot, ts = sum(inst.values()), max(sum(samp.values()), 1)
    text = open(src_file).read().splitlines() if src_file and os.path.exists(src_file) else []
    print(f"# {kern}: {tot} warp instructions; by source line of {src_file}")
    for ln, n in inst.most_common(top):
        code = text[ln - 1].strip()[:100] if ln and ln <= len(text) else ""
        print(f"{ln:5d} {100*n/tot:5.1f}% inst {100*samp[ln]/ts:5.1f}% samples | {code}")


if __name__ == "__main__":
    main()
