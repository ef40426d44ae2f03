#!/usr/bin/env python
"""Issue / register-bandwidth model of a SASS loop (B200, sm_100a).

    cuobjdump -sass -fun <mangled kernel> lib.so > k.sass
    python profiles/sass_model.py k.sass 0x18c0 0x2fe0 [--skip 0x2c50:0x2df0 ...]

Per instruction between the two addresses (inclusive):
  * issue slots: 1;
  * FMA-pipe cycles: 2 for packed fp32x2 (FFMA2/FMUL2/FADD2), 1 for scalar FFMA/FMUL/FADD/IMAD/HFMA2;
  * register-file reads: distinct 32-bit vector registers among the SOURCE operands (a .F32x2 / 64-bit operand is two
    registers, a ".F32" broadcast operand one), not counting an operand that the previous instruction left in the
    same slot's reuse cache (".reuse") and not counting RZ / uniform registers / immediates / constant-bank operands.
    Measured rule (profiles/microbench/regbw*.cu): the register file of an SM sub-partition feeds ~2 such reads per clock.
The loop's cost per warp and sub-partition is then max(issue slots, sum over instructions of max(pipe cycles, reads/2))
for the FMA-pipe instructions plus reads/2 for the rest, and MUFU needs 8 cycles per warp instruction (4 lanes/clk).
"""
import re
import sys

PACKED = {"FFMA2", "FMUL2", "FADD2"}
FMA_SCALAR = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FMNMX3"}
NO_DEST = {"STG", "STS", "ST", "BRA", "EXIT", "BAR", "BSSY", "BSYNC", "CALL", "RET", "NOP", "WARPSYNC", "LDGSTS", "DEPBAR",
           "MEMBAR", "RED", "STL", "ATOMS", "ERRBAR", "CCTL", "UBLKCP", "SYNCS", "NANOSLEEP", "YIELD", "FENCE"}
WIDE64 = {"DFMA", "DADD", "DMUL", "F2F.F64.F32", "DSETP"}

line_re = re.compile(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);")


def parse(path):
    out = []
    for ln in open(path):
        m = line_re.match(ln)
        if not m:
            continue
        out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def split_ops(text):
    # "@P0 FFMA2 R4, R4.F32x2.HI_LO, R2.F32x2.HI_LO, UR6.F32" -> pred, opcode, [operands]
    pred = None
    if text.startswith("@"):
        pred, text = text.split(None, 1)
    parts = text.split(None, 1)
    opcode = parts[0]
    ops = []
    if len(parts) > 1:
        depth = 0
        cur = ""
        for ch in parts[1]:
            if ch == "[":
                depth += 1
            if ch == "]":
                depth -= 1
            if ch == "," and depth == 0:
                ops.append(cur.strip())
                cur = ""
            else:
                cur += ch
        if cur.strip():
            ops.append(cur.strip())
    return pred, opcode, ops


reg_re = re.compile(r"(?<![UA-Za-z_])R(\d+)")


def operand_regs(op, opcode_base, opcode_full):
    """32-bit vector registers read by one source operand."""
    if "[" in op:  # address operand: R or R.64 (+ UR) -> count base register pair for 64-bit addresses
        regs = reg_re.findall(op)
        out = []
        for r in regs:
            r = int(r)
            out += [r, r + 1] if ".64" in op else [r]
        return out
    m = reg_re.search(op)
    if not m or "RZ" == op.strip("-|~!"):
        return []
    r = int(m.group(1))
    if ".F32x2" in op or ".64" in op:
        return [r, r + 1]
    if opcode_base in PACKED and ".F32" not in op:
        return [r, r + 1]
    if opcode_base in ("DFMA", "DADD", "DMUL", "DSETP"):
        return [r, r + 1]
    return [r]


def analyse(instrs, lo, hi, skips):
    issue = 0
    fma_cycles = 0.0
    other_rf = 0.0
    mufu = 0
    counts = {}
    reads_total = 0
    prev_reuse = {}
    detail = {"packed3": 0, "packed2": 0, "scalar3": 0, "scalar2": 0}
    for addr, text in instrs:
        if addr < lo or addr > hi:
            continue
        if any(a <= addr <= b for a, b in skips):
            continue
        pred, opcode, ops = split_ops(text)
        base = opcode.split(".")[0]
        counts[base] = counts.get(base, 0) + 1
        issue += 1
        srcs = ops if base in NO_DEST else ops[1:]
        if base in ("ISETP", "FSETP", "DSETP", "PLOP3", "UISETP"):
            srcs = ops[2:]  # two predicate destinations
        regs = set()
        new_reuse = {}
        for slot, op in enumerate(srcs):
            rr = operand_regs(op, base, opcode)
            if not rr:
                continue
            key = (slot, tuple(rr))
            if prev_reuse.get(slot) == tuple(rr):
                pass  # served by the reuse cache
            else:
                regs.update(rr)
            if ".reuse" in op:
                new_reuse[slot] = tuple(rr)
        prev_reuse = new_reuse
        nreads = len(regs)
        reads_total += nreads
        if base == "MUFU":
            mufu += 1
        if base in PACKED:
            c = max(2.0, nreads / 2.0)
            fma_cycles += c
            detail["packed3" if nreads > 4 else "packed2"] += 1
        elif base in FMA_SCALAR:
            c = max(1.0, nreads / 2.0)
            fma_cycles += c
            detail["scalar3" if nreads > 2 else "scalar2"] += 1
        else:
            other_rf += nreads / 2.0
    return dict(issue=issue, fma_cycles=fma_cycles, other_rf=other_rf, mufu=mufu, counts=counts, reads=reads_total, detail=detail)


def main():
    path = sys.argv[1]
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
    skips = []
    if "--skip" in sys.argv:
        for s in sys.argv[sys.argv.index("--skip") + 1:]:
            a, b = s.split(":")
            skips.append((int(a, 16), int(b, 16)))
    r = analyse(parse(path), lo, hi, skips)
    top = sorted(r["counts"].items(), key=lambda kv: -kv[1])
    print("instructions:", r["issue"], " ".join(f"{k}:{v}" for k, v in top[:24]))
    print("forms:", r["detail"], " RF reads:", r["reads"])
    rf_total = r["reads"] / 2.0
    print(f"per loop trip and warp: issue {r['issue']}  FMA-pipe/RF cycles {r['fma_cycles']:.0f} (+{r['other_rf']:.0f} RF cycles of other instructions)"
          f"  RF-only bound {rf_total:.0f}  MUFU cycles {8 * r['mufu']}")
    bound = max(r["issue"], r["fma_cycles"] + r["other_rf"], 8 * r["mufu"])
    print(f"bound per trip: {bound:.0f} cycles")


if __name__ == "__main__":
    main()
