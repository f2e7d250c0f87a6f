#!/usr/bin/env python
"""Commit-able SASS evidence: `cuobjdump -sass` of the hot kernels of libdogstep.so -> profiles/<tag>_sass_<kernel>.txt plus a
per-kernel opcode histogram (profiles/<tag>_sass_summary.json).  Runs in the build container (no GPU needed).

    python scripts/dump_sass.py r2            # after python __graft_entry__.py
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "exploring-muzero-on-dog_b200", "libdogstep.so")
HOT = {  # file tag -> substring of the mangled name
    "k_madn_det_play_cta": "k_madn_det_play_ctaILj211EE",
    "k_dog_play_random": "k_dog_play_randomE",
    "k_mcts_expand_select_8_2": "k_mcts_expand_selectILi8ELi2EE",
    "k_mcts_expand_select_10_1": "k_mcts_expand_selectILi10ELi1EE",
    "k_replay_save": "k_replay_saveE",
    "k_ttt_search": "k_ttt_searchILb1EE",
    "k_madn_det_random_step": "k_madn_det_random_stepE",
}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    blocks = re.split(r"(?=^\s+Function : )", txt, flags=re.M)
    summary = {"arch": arch, "kernels": {}}
    for name, needle in HOT.items():
        blk = next((b for b in blocks if needle in b.split("\n", 1)[0]), None)
        if blk is None:
            print("missing", name)
            continue
        path = os.path.join(ROOT, "profiles", f"{tag}_sass_{name}.txt")
        # keep address + instruction; the 128-bit encodings (two hex words per instruction) only double the file
        lean = re.sub(r"[ \t]+/\* 0x[0-9a-f]{16} \*/[ \t]*$", "", blk, flags=re.M)
        lean = "\n".join(line.rstrip() for line in lean.split("\n") if line.strip())
        with open(path, "w") as f:
            f.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({', '.join(arch)}); instruction encodings stripped\n" + lean + "\n")
        ops = collections.Counter(m.group(1).split(".")[0] for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk, flags=re.M))
        summary["kernels"][name] = {"instructions": sum(ops.values()), "top_opcodes": dict(ops.most_common(16)),
                                    "tensor_or_tma": sorted(k for k in ops if k.startswith(("UTC", "UTMA", "UBLK", "HMMA", "QMMA")))}
        print(name, sum(ops.values()), "SASS instructions;", ", ".join(f"{k} {v}" for k, v in ops.most_common(8)))
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
