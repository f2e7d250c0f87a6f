"""Per-function / per-phase instruction budget of one kernel from an ncu report captured with --import-source on:
   python profiles/phase_budget.py <report.ncu-rep> <libdogstep.so> <kernel substring> <units> [out.json]
Joins `ncu --page source --csv` (per-SASS-instruction executed counts, in program order) with `nvdisasm -g` line info, maps every
source line to the function that contains it (the device functions are all inlined into the kernel, -lineinfo keeps their
lines) and groups the functions into phases.  `units` = what one launch processed (env steps): the budget is printed per unit."""
import csv
import json
import os
import re
import subprocess
import sys
import tempfile

PHASES = [  # (phase, regex on the function name), first match wins
    ("threefry / key derivation", r"threefry|split_i|bits_i|uniform_i|randint_i|bits_to"),
    ("categorical draw", r"categorical"),
    ("deal (distribute_cards)", r"deal|distribute|deck|shuffle"),
    ("legal mask", r"mask|val_|valid|canonical|hot7_ok|seven_ok|path_|can_"),
    ("transition (env_step / no_step)", r"env_step|no_step|step_|apply|winner|next_player|swap_phase|play_phase"),
    ("load / store / game queue", r"load|store|next_game"),
    ("kernel loop, barriers, task queue", r"^k_"),
]


def functions_of(path):
    """[(first line, name)] of the function definitions of a source file (brace-less heuristics are enough for this code base)"""
    out = []
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+|inline\s+|__global__\s+|__device__\s+|__host__\s+|__forceinline__\s+|DS_FN\s+|"
                     r"__launch_bounds__\([^)]*\)\s*)*[\w:<>,\s\*&]+?\b(\w+)\s*\([^;]*$")
    for i, line in enumerate(open(path).read().splitlines(), 1):
        if line.startswith((" ", "\t", "//", "#", "}")) and not re.match(r"^\s*(DS_FN|__device__|__global__|template)", line):
            continue
        m = pat.match(line)
        if m and m.group(1) not in ("if", "for", "while", "switch", "return", "sizeof", "defined"):
            out.append((i, m.group(1)))
    return out


def main(rep, so, kern, units, out_json=None):
    units = float(units)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ii = hdr.index("Instructions Executed")
    ti = hdr.index("Thread Instructions Executed") if "Thread Instructions Executed" in hdr else None
    si = hdr.index("# Samples") if "# Samples" in hdr else None
    counts = [(int(r[ii]), int(r[ti]) if ti is not None else 0, int(r[si]) if si is not None else 0)
              for r in rows[h + 1:] if len(r) > ii and r[0].startswith("0x")]
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
    assert lines, "kernel not found in the library"
    counts = counts[:len(lines)]
    csrc = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc")
    fn_tables = {}
    per_fn = {}
    for (n, tn, s), (f, l) in zip(counts, lines):
        if f not in fn_tables:
            p = os.path.join(csrc, f)
            fn_tables[f] = functions_of(p) if os.path.exists(p) else []
        name = "?"
        for first, nm in fn_tables[f]:
            if first <= l:
                name = nm
            else:
                break
        a = per_fn.setdefault((f, name), [0, 0, 0, 0])
        a[0] += n; a[1] += tn; a[2] += s; a[3] += 1
    tot = sum(a[0] for a in per_fn.values())
    tot_s = max(1, sum(a[2] for a in per_fn.values()))
    phases = {}
    for (f, name), a in per_fn.items():
        ph = next((p for p, rx in PHASES if re.search(rx, name)), "other")
        b = phases.setdefault(ph, [0, 0, 0])
        b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
    print(f"{kern}: {len(counts)} SASS instructions, {tot} warp-instructions executed = {tot / units:.1f} per unit ({units:.0f} units)")
    print(f"{'phase':42s} {'warp-inst/unit':>14s} {'share':>7s} {'lanes':>6s} {'samples':>8s}")
    res = {"kernel": kern, "units": units, "warp_instructions": tot, "warp_instructions_per_unit": tot / units, "phases": {}, "functions": {}}
    for ph, b in sorted(phases.items(), key=lambda kv: -kv[1][0]):
        lanes = b[1] / b[0] if b[0] else 0
        print(f"{ph:42s} {b[0] / units:14.1f} {100 * b[0] / tot:6.1f}% {lanes:6.1f} {100 * b[2] / tot_s:7.1f}%")
        res["phases"][ph] = {"warp_inst_per_unit": b[0] / units, "share": b[0] / tot, "active_lanes": lanes, "stall_sample_share": b[2] / tot_s}
    print()
    for (f, name), a in sorted(per_fn.items(), key=lambda kv: -kv[1][0])[:28]:
        print(f"  {a[0] / units:9.1f} /unit {100 * a[0] / tot:5.1f}%  lanes {a[1] / max(a[0], 1):5.1f}  smp {100 * a[2] / tot_s:5.1f}%  sass {a[3]:5d}  {f}:{name}")
        res["functions"][f"{f}:{name}"] = {"warp_inst_per_unit": a[0] / units, "share": a[0] / tot, "active_lanes": a[1] / max(a[0], 1)}
    if out_json:
        json.dump(res, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:6])
