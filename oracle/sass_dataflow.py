#!/usr/bin/env python
"""Symbolic dataflow extractor for fp32 SASS (TEST INFRASTRUCTURE ONLY).

Purpose: the reference rasterizer is compiled with nvcc's default -fmad=true, so
*both* NVVM and ptxas contract mul+add pairs into FFMA.  Which pairs are fused
decides the last bit of radii / tile rectangles / depths, i.e. the integer
outputs that must match bit-exactly.  This tool walks the SASS of one kernel of
the compiled reference (oracle/_ref/obj_*/forward.cu.o) in program order,
tracks every fp32 register as an expression tree, recognises the IEEE
div/rcp/sqrt expansions, and prints the expression stored by every STG.  The
printed trees are the arithmetic specification that oracle/ols_oracle.cpp and
the CUDA kernels restate with explicit fmaf()/__fmaf_rn().

Usage: python oracle/sass_dataflow.py <object-or-cubin> <kernel-substring> [--from 0xADDR --to 0xADDR]
"""
import re
import subprocess
import sys

LINE = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(?:(@!?U?P[0-9T])\s+)?([A-Z0-9_.]+)\s*(.*?)\s*;")


def parse_operand(tok):
    tok = tok.strip()
    neg = absv = False
    if tok.startswith("-"):
        neg, tok = True, tok[1:]
    if tok.startswith("|") and tok.endswith("|"):
        absv, tok = True, tok[1:-1]
    tok = tok.replace(".reuse", "")
    return neg, absv, tok


class Machine:
    def __init__(self):
        self.regs = {}
        self.nleaf = 0
        self.stores = []

    def get(self, tok):
        neg, absv, name = parse_operand(tok)
        if name == "RZ" or name == "URZ":
            e = ("const", "0")
        elif re.fullmatch(r"U?R\d+", name):
            e = self.regs.get(name, ("reg", name))
        elif name.startswith("c[") or name.startswith("desc["):
            e = ("cmem", name)
        else:
            e = ("const", name)
        if absv:
            e = ("abs", e)
        if neg:
            e = ("neg", e)
        return e

    def step(self, addr, pred, op, args):
        toks = [a.strip() for a in split_args(args)]
        base = op.split(".")[0]
        if base in ("FFMA", "FMUL", "FADD", "FMNMX", "DFMA", "DMUL", "DADD"):
            d = toks[0]
            srcs = [self.get(t) for t in toks[1:] if not re.fullmatch(r"!?U?PT|!?U?P\d", t.strip())]
            mods = ".".join(m for m in op.split(".")[1:] if m not in ("FTZ",))
            name = base.lower() + (("." + mods.lower()) if mods else "")
            if base == "FMNMX":
                # last arg is predicate: PT -> min, !PT -> max
                p = toks[-1]
                name = "min" if p == "PT" else ("max" if p == "!PT" else "minmax[%s]" % p)
            self.regs[d] = (name,) + tuple(srcs)
        elif base == "MUFU":
            self.regs[toks[0]] = (op.lower(), self.get(toks[1]))
        elif base in ("MOV", "UMOV"):
            self.regs[toks[0]] = self.get(toks[1])
        elif op.startswith("IMAD.MOV"):
            self.regs[toks[0]] = self.get(toks[3])
        elif base == "HFMA2" and toks[1].replace("-", "") == "RZ" and toks[2] == "RZ":
            # HFMA2 Rd, -RZ, RZ, hi, lo : materialises an fp32 immediate from two halves
            self.regs[toks[0]] = ("const", "h2(%s,%s)" % (toks[3], toks[4]))
        elif base in ("LDG", "LDC", "LDCU", "ULDC", "LD", "LDS"):
            self.nleaf += 1
            src = toks[1] if len(toks) > 1 else "?"
            width = 2 if ".64" in op else (4 if ".128" in op else 1)
            m = re.fullmatch(r"(U?R)(\d+)", toks[0])
            for i in range(width):
                nm = "%s%d" % (m.group(1), int(m.group(2)) + i) if m else toks[0]
                self.regs[nm] = ("load", "%s@%s%s" % (src, addr, "" if width == 1 else "+%d" % (4 * i)))
        elif base == "STG":
            width = 2 if ".64" in op else (4 if ".128" in op else 1)
            m = re.fullmatch(r"(U?R)(\d+)", toks[1].replace(".reuse", ""))
            vals = []
            for i in range(width):
                nm = "%s%d" % (m.group(1), int(m.group(2)) + i) if m else toks[1]
                vals.append(self.regs.get(nm, ("reg", nm)) if nm not in ("RZ",) else ("const", "0"))
            self.stores.append((addr, op, toks[0], vals))
        elif base in ("FSEL", "SEL"):
            self.regs[toks[0]] = ("sel[%s]" % toks[3], self.get(toks[1]), self.get(toks[2]))
        elif base in ("F2F", "F2I", "I2F", "I2FP", "FRND", "F2FP"):
            self.regs[toks[0]] = (op.lower(), self.get(toks[-1]))
        elif base in ("FSETP", "FCHK", "ISETP", "PLOP3", "BRA", "BSSY", "BSYNC", "CALL", "EXIT", "NOP",
                      "RET", "UISETP", "LEPC", "S2R", "S2UR", "BAR", "DSETP"):
            if base == "S2R":
                self.regs[toks[0]] = ("sreg", toks[1])
        else:
            if toks and re.fullmatch(r"U?R\d+", toks[0]):
                self.regs[toks[0]] = ("int", "%s@%s" % (op, addr))


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "[(":
            depth += 1
        elif ch in "])":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def is_const(e, v):
    return e[0] == "const" and e[1] in v


def neg_of(e):
    return e[1] if e[0] == "neg" else ("neg", e)


def simplify(e, memo):
    """Recognise the IEEE expansions emitted by ptxas and rewrite them as div/rcp/sqrt nodes."""
    key = id(e)
    if key in memo:
        return memo[key]
    if not isinstance(e, tuple) or e[0] in ("const", "load", "cmem", "reg", "int", "sreg"):
        memo[key] = e
        return e
    e = (e[0],) + tuple(simplify(x, memo) if isinstance(x, tuple) else x for x in e[1:])
    r = e
    # rcp.rn(b): r0=rcp(b); e=ffma(b,r0,-1); e'=fadd(-e,-0); r=ffma(r0,e',r0)
    if e[0] == "ffma" and len(e) == 4 and e[1] == e[3] and e[1][0] == "mufu.rcp":
        r = ("RCP", e[1][1])
    # newton refinement used by div: r1 = ffma(r0, ffma(-b, r0, 1), r0)
    if e[0] == "ffma" and len(e) == 4 and e[1][0] == "mufu.rcp" and e[3] == e[1] and e[2][0] == "ffma":
        r = ("RCP", e[1][1])
    # div.rn(a,b): q0=ffma(a,r1,0)|fmul ; rem=ffma(-b,q0,a); q=ffma(r1,rem,q0)
    if e[0] == "ffma" and len(e) == 4 and e[1][0] == "RCP" and e[2][0] == "ffma" and e[3][0] in ("ffma", "fmul"):
        b = e[1][1]
        q0 = e[3]
        a = None
        if q0[0] == "ffma" and is_const(q0[3], ("0",)):
            a = q0[1] if q0[2] == e[1] else (q0[2] if q0[1] == e[1] else None)
        elif q0[0] == "fmul":
            a = q0[1] if q0[2] == e[1] else (q0[2] if q0[1] == e[1] else None)
        if a is not None:
            r = ("DIV", a, b)
    # sqrt.rn(x): y=rsq(x); g=fmul(x,y); h=fmul(y,.5); r=ffma(-g,g,x); res=ffma(r,h,g)
    if e[0] == "ffma" and len(e) == 4 and e[3][0] in ("fmul", "fmul.ftz") and e[2][0] in ("fmul", "fmul.ftz") \
            and e[1][0] == "ffma":
        g = e[3]
        if any(isinstance(t, tuple) and t[0] == "mufu.rsq" for t in g[1:]):
            x = [t for t in g[1:] if not (isinstance(t, tuple) and t[0] == "mufu.rsq")]
            if x:
                r = ("SQRT", x[0])
    memo[key] = r
    return r


def fmt(e, names, depth=0):
    if e[0] == "const":
        return e[1]
    if e[0] in ("load", "cmem", "reg", "int", "sreg"):
        return "%s<%s>" % (e[0], e[1])
    if e[0] == "neg":
        return "-" + fmt(e[1], names, depth + 1)
    if e[0] == "abs":
        return "|" + fmt(e[1], names, depth + 1) + "|"
    k = id(e)
    if k in names:
        return names[k]
    return "%s(%s)" % (e[0], ", ".join(fmt(x, names, depth + 1) for x in e[1:]))


def count_uses(e, uses, seen):
    if not isinstance(e, tuple) or e[0] in ("const", "load", "cmem", "reg", "int", "sreg"):
        return
    uses[id(e)] = uses.get(id(e), 0) + 1
    if id(e) in seen:
        return
    seen[id(e)] = e
    for x in e[1:]:
        if isinstance(x, tuple):
            count_uses(x, uses, seen)


def main():
    obj, kern = sys.argv[1], sys.argv[2]
    lo = hi = None
    if "--from" in sys.argv:
        lo = int(sys.argv[sys.argv.index("--from") + 1], 16)
    if "--to" in sys.argv:
        hi = int(sys.argv[sys.argv.index("--to") + 1], 16)
    take = set()
    for i, a in enumerate(sys.argv):
        if a == "--take":
            take.add(int(sys.argv[i + 1], 16))
    sass = subprocess.check_output(["cuobjdump", "-sass", obj], text=True).splitlines()
    active, m, prog = False, Machine(), []
    for ln in sass:
        if "Function :" in ln:
            active = kern in ln
            continue
        if not active:
            continue
        mt = LINE.match(ln)
        if mt:
            prog.append((int(mt.group(1), 16), mt.group(2), mt.group(3), mt.group(4)))
    index = {a: i for i, (a, _, _, _) in enumerate(prog)}
    pc = index[lo] if lo is not None else 0
    steps = 0
    while pc < len(prog) and steps < 100000:
        steps += 1
        addr, pred, op, args = prog[pc]
        if hi is not None and addr > hi:
            break
        if op.split(".")[0] in ("EXIT", "RET") and pred is None:
            break
        if op.split(".")[0] == "BRA":
            toks = [t.strip() for t in split_args(args)]
            tgt = int(toks[-1], 16)
            cond = pred is not None or len(toks) > 1
            if tgt > addr and tgt in index:
                between = prog[pc + 1:index[tgt]]
                guards_call = any(o.startswith("CALL") for _, _, o, _ in between)
                if (not cond) or guards_call or addr in take:
                    pc = index[tgt]
                    continue
            pc += 1
            continue
        m.step("0x%04x" % addr, pred, op, args)
        pc += 1
    memo = {}
    stores = [(a, op, dst, [simplify(v, memo) for v in vals]) for a, op, dst, vals in m.stores]
    uses, seen = {}, {}
    for _, _, _, vals in stores:
        for v in vals:
            count_uses(v, uses, seen)
    # name shared sub-expressions t0, t1, ... in first-use (post-)order
    names, order = {}, []

    def visit(e):
        if not isinstance(e, tuple) or e[0] in ("const", "load", "cmem", "reg", "int", "sreg"):
            return
        for x in e[1:]:
            if isinstance(x, tuple):
                visit(x)
        if uses.get(id(e), 0) > 1 and id(e) not in names and e[0] not in ("neg", "abs"):
            names[id(e)] = None
            order.append(e)

    for _, _, _, vals in stores:
        for v in vals:
            visit(v)
    for i, e in enumerate(order):
        tmp = dict((k, v) for k, v in names.items() if v is not None)
        s = "%s(%s)" % (e[0], ", ".join(fmt(x, tmp) for x in e[1:]))
        names[id(e)] = "t%d" % i
        print("t%d = %s" % (i, s))
    for a, op, dst, vals in stores:
        for i, v in enumerate(vals):
            print("STORE %s %s %s [%d] = %s" % (a, op, dst, i, fmt(v, names)))


if __name__ == "__main__":
    main()
