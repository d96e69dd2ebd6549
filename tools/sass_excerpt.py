#!/usr/bin/env python
"""SASS of the scan kernels' hot loops from the built library, for profiles/: per kernel the opcode histogram and the
instruction listing (encodings stripped).  Usage: python tools/sass_excerpt.py [substring of the mangled name ...]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "smart_b200", "lib", "libsmartb200.so")
LIST = 140  # instructions listed per kernel
DEFAULT = ["scan_block_kernelILi4ELi4ELi4E", "scan_block_kernelILi2ELi3ELi4E", "scan_window_kernelILi3ELi4ELb0ELi4E",
           "scan_pswar_kernelILi1ELi12ELb1E", "scan_pswar_kernelILi2ELi10ELb0E"]


def main():
    wanted = sys.argv[1:] or DEFAULT
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    for w in wanted:
        for f in funcs[1:]:
            name, _, body = f.partition("\n")
            if w not in name:
                continue
            demangled = subprocess.run(["cu++filt", name.strip()], capture_output=True, text=True).stdout.strip() or name
            ins = re.findall(r"/\*([0-9a-f]{4})\*/\s+(.*?);", body)
            # the hot loop: from the first streaming load to the backward branch that closes the group loop
            first = next((i for i, (_, t) in enumerate(ins) if "LDG.E.NA" in t or "LDG.E.ENL2" in t or "LDG.E.256" in t), 0)
            last = first
            for i, (addr, t) in enumerate(ins):
                m = re.search(r"BRA\s+(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", t)
                if m and m.group(1) and int(m.group(1), 16) <= int(ins[first][0], 16) and i > first:
                    last = i
            loop = ins[first:last + 1]
            hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in loop)
            print(f"## `{demangled}`\n")
            print(f"{len(loop)} instructions from the first streaming load to the loop's closing branch "
                  f"(one warp iteration = 4 KiB of input per warp); rare-path calls are outside this range or predicated off.\n")
            print("| opcode | count |\n|---|---|")
            for op, c in hist.most_common():
                print(f"| {op} | {c} |")
            print("\n```")
            for addr, t in loop[:LIST]:
                print(f"/*{addr}*/  {t}")
            if len(loop) > LIST:
                print(f"... ({len(loop) - LIST} more instructions of the same unrolled pattern)")
            print("```\n")
            break


if __name__ == "__main__":
    main()
