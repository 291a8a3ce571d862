#!/usr/bin/env python
"""Static instruction mix of the loops of one kernel (no GPU needed): disassembles a library / object with
`cuobjdump -sass`, finds the backward branches of the first function whose mangled name matches REGEX and prints, per
loop, the instruction count and how many of them are shared loads, fp64 operations, global stores / loads, moves, selects
and synchronisation.  This is how DESIGN.md section 10 counts the 3D interior loop of the apply kernel (70 instructions
per two nodes) and checks that a build variant leaves the default kernel's SASS unchanged (`--dump` + diff).

    python tools/sass_loops.py [FILE] [REGEX] [--min-lds N] [--dump]
    python tools/sass_loops.py homogenization.jl_b200/libhmg_b200.so 'apply_kernelILi3ELi32ELi0ELb0ELb0' --min-lds 6
"""
import re
import subprocess
import sys


def function_sass(path, pattern):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout.splitlines()
    rx = re.compile(pattern)
    body, inside, name = [], False, None
    for line in out:
        m = re.search(r"Function : (\S+)", line)
        if m:
            if inside:
                break
            if rx.search(m.group(1)):
                inside, name = True, m.group(1)
            continue
        if inside:
            m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                body.append((int(m.group(1), 16), m.group(2).strip()))
    return name, body


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    path = args[0] if args else "homogenization.jl_b200/libhmg_b200.so"
    pattern = args[1] if len(args) > 1 else "apply_kernelILi3ELi32ELi0ELb0ELb0"
    min_lds = int(sys.argv[sys.argv.index("--min-lds") + 1]) if "--min-lds" in sys.argv else 0
    name, ins = function_sass(path, pattern)
    if not ins:
        raise SystemExit(f"no function matching {pattern!r} in {path}")
    if "--dump" in sys.argv:                       # addresses stripped: two builds can be diffed
        for _, t in ins:
            print(t)
        return
    print(f"{name}: {len(ins)} instructions")
    classes = (("LDS", r"\bLDS"), ("fp64", r"\bD(FMA|ADD|MUL)\b"), ("STG", r"\bSTG"), ("LDG", r"\bLDG"),
               ("MOV", r"MOV"), ("FSEL", r"\bFSEL"), ("sync", r"SYNCS|\bBAR\b|UBLKCP"))
    for addr, text in ins:
        if "BRA" not in text:
            continue
        m = re.search(r"0x([0-9a-f]+)", text)
        if not m or int(m.group(1), 16) >= addr:
            continue
        start = int(m.group(1), 16)
        body = [t for a, t in ins if start <= a <= addr]
        count = {k: sum(1 for t in body if re.search(rx, t)) for k, rx in classes}
        if count["LDS"] < min_lds:
            continue
        other = len(body) - sum(count.values())
        print(f"  loop {start:#07x}-{addr:#07x}: {len(body):5d} instr  " + "  ".join(f"{k} {v}" for k, v in count.items())
              + f"  other {other}")


if __name__ == "__main__":
    main()
