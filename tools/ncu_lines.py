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
    dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines_of_inst = []
    cur_line, active = None, False
    fresh, cur_is_cu = True, False
    src_file = None
    for l in dis.splitlines():
        if l.startswith(".text."):
            active = kern in l
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            # -gi prints the inline chain innermost first: keep the innermost frame that lies in a .cu file (intrinsics
            # from the toolkit headers are attributed to the line that calls them)
            if fresh or not cur_is_cu:
                if fresh or m.group(1).endswith(".cu"):
                    cur_line = int(m.group(2))
                    cur_is_cu = m.group(1).endswith(".cu")
                    if cur_is_cu:
                        src_file = m.group(1)
            fresh = False
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines_of_inst.append(cur_line)
            fresh = True
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # a report may hold several launches: sections start with a "Kernel Name" row followed by the header row
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    want = re.sub(r"^_Z\d*", "", kern)
    base = re.match(r"[A-Za-z_]+", want).group(0)
    pick = [i for i in starts if base in rows[i][1] and ((base + "<") in rows[i][1]) == ("IL" in want)]
    s0 = (pick or starts)[0]
    s1 = min([i for i in starts if i > s0] + [len(rows)])
    h = rows[s0 + 1]
    data = [r for r in rows[s0 + 2:s1] if len(r) > h.index('Instructions Executed')]
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
