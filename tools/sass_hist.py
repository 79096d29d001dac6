"""Opcode histogram of the hot loop of every Smith-Waterman kernel in libswb200.so (static evidence beside ncu).

    python tools/sass_hist.py [path/to/libswb200.so] > profiles/rNN/sass_hot_loops.json

For each `sw_kernel*` function: the smallest loop of at least 800 instructions is the steady 16-step body (the
`#pragma unroll 1` loop of sw_two_pairs); its instructions are counted by opcode and by issue pipe (ALU pipe: PRMT, VIMNMX*, VIADDMNMX,
LOP3, SHF, IADD3, ISETP, SEL; FMA pipe: IMAD*, VIADD, HFMA2, HADD2 -- profiles/INT_PEAK.json).  Needs cuobjdump."""
import collections
import json
import os
import re
import subprocess
import sys

ALU = ("PRMT", "VIMNMX", "VIMNMX3", "VIADDMNMX", "LOP3", "SHF", "IADD3", "ISETP", "SEL", "LEA", "HMNMX2", "FMNMX")
FMA = ("IMAD", "VIADD", "HFMA2", "HADD2", "FFMA", "FMUL", "FADD")
LSU = ("LDS", "STS", "LDG", "STG", "LDL", "STL")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "smith-waterman-simd_b200", "libswb200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    out = {}
    for chunk in re.split(r"\n\s+Function : ", sass)[1:]:
        name = chunk.split("\n", 1)[0].strip()
        if "sw_kernel" not in name:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)\s*([^;]*);", chunk)
        # every backward branch closes a loop; the steady 16-step body is the SMALLEST loop that still holds at least
        # 800 instructions (16 steps x ~67): the strip loop and a persistent kernel's item loop are larger, nothing else is as big
        loops = []
        for addr, op, args in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", args)
                if m and int(m.group(1), 16) < int(addr, 16):
                    loops.append((int(addr, 16) - int(m.group(1), 16), int(m.group(1), 16), int(addr, 16)))
        loops = [l for l in loops if l[0] >= 800 * 16]
        if not loops:
            continue
        lo, hi = min(loops)[1:]
        hist = collections.Counter(op.split(".")[0] for addr, op, _ in ins if lo <= int(addr, 16) <= hi)
        total = sum(hist.values())
        pipe = {"alu": sum(v for k, v in hist.items() if k in ALU), "fma": sum(v for k, v in hist.items() if k in FMA),
                "lsu": sum(v for k, v in hist.items() if k in LSU)}
        out[name] = {"steady_body_instructions": total, "per_step": round(total / 16.0, 2), "alu_pipe_per_step": round(pipe["alu"] / 16.0, 2),
                     "alu_pipe_per_cell": round(pipe["alu"] / 16.0 / 32.0, 3), "fma_pipe_per_step": round(pipe["fma"] / 16.0, 2),
                     "lsu_per_step": round(pipe["lsu"] / 16.0, 2), "opcodes": dict(hist.most_common())}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
