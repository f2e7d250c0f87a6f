"""Per-source-line instruction / stall-sample shares of one kernel:
   python profiles/hotlines.py <report.ncu-rep> <libdogstep.so> <kernel substring> [top]
Joins `ncu --page source --csv` (per-SASS-instruction counts, in program order) with `nvdisasm -g` line info."""
import csv
import os
import re
import subprocess
import sys
import tempfile


def main(rep, so, kern, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    ni = hdr.index("stall_no_inst") if "stall_no_inst" in hdr else None
    counts = []
    for r in rows[h + 1:]:
        if len(r) > ii and r[0].startswith("0x"):
            counts.append((int(r[ii]), int(r[si]), int(r[ni]) if ni is not None else 0))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    lines = None
    for f in sorted(os.listdir(tmp)):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kern not in out:
            continue
        sec, cur, lines = False, ("?", 0), []
        for ln in out.splitlines():
            if ln.startswith(".text."):
                sec = kern in ln
                continue
            if not sec:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                lines.append(cur)
        break
    assert lines, 'kernel not found in the library'
    counts = counts[:len(lines)]  # several launches in one report: keep the first
    agg = {}
    for (n, s, ns), key in zip(counts, lines):
        a = agg.setdefault(key, [0, 0, 0, 0])
        a[0] += n; a[1] += s; a[2] += ns; a[3] += 1
    tot_i, tot_s = sum(a[0] for a in agg.values()), max(1, sum(a[1] for a in agg.values()))
    print(f"{len(counts)} SASS instructions, {tot_i} warp-instructions executed, {tot_s} samples")
    srcs = {}
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        if f not in srcs:
            p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        print(f"{a[1] / tot_s * 100:5.1f}% smp {a[0] / tot_i * 100:5.1f}% inst  no_inst {a[2]:6d}  sass {a[3]:4d}  {f}:{l}  {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
