#!/usr/bin/env python
"""Pruned SSA listing of the floating-point dataflow behind the global stores of one SASS kernel.

ptxas is allowed to fuse PTX `mul.f32` + `add.f32` (no rounding suffix) into FFMA, so the final word
on which products of the reference are fused is its SASS, not its PTX.  This tool linearly scans a
`cuobjdump -sass` listing, tracks the current symbolic value of every register (turning the register
machine into SSA), and prints, for each STG, the tree of FP operations that produced the stored value.
Control flow is ignored (exact for the straight-line per-Gaussian kernels; loops show as re-definitions).

usage: sass_expr.py file.sass kernel_substring [--param-base 0x380]
"""
import re
import sys

FP_OPS = ("FMUL", "FADD", "FFMA", "MUFU", "DFMA", "DMUL", "DADD", "F2F", "FMNMX", "FSEL", "I2F", "F2I", "FRND",
          "I2FP", "FSETP", "DSETP", "DMNMX", "FCHK", "FSET")


def kernel_lines(text, key):
    out, on = [], False
    for line in text.split("\n"):
        if "Function :" in line:
            on = key in line
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                out.append((m.group(1), m.group(2).strip()))
    if not out:
        raise SystemExit(f"kernel containing {key!r} not found")
    return out


def main():
    path, key = sys.argv[1], sys.argv[2]
    pbase = 0x380
    if "--param-base" in sys.argv:
        pbase = int(sys.argv[sys.argv.index("--param-base") + 1], 16)
    lines = kernel_lines(open(path).read(), key)
    val = {}      # reg -> node name / leaf
    tag = {}      # reg -> param tag (for address provenance)
    nodes = {}    # name -> text
    order = []
    used_by_store = []

    def new_node(text):
        n = f"t{len(order)}"
        nodes[n] = text
        order.append(n)
        return n

    def src(tok):
        tok = tok.strip()
        neg = tok.startswith("-")
        ab = "|" in tok
        t = tok.lstrip("-").replace("|", "")
        t = re.sub(r"\.(reuse|H0_H0|H1_H1)$", "", t)
        t = t.replace(".reuse", "")
        if re.fullmatch(r"R\d+", t):
            v = val.get(t, t)
        elif t == "RZ":
            v = "0"
        elif t.startswith("c["):
            m = re.match(r"c\[0x0\]\[0x([0-9a-f]+)\]", t)
            v = f"P@{int(m.group(1), 16) - pbase}" if m else t
        elif re.fullmatch(r"UR\d+", t):
            v = val.get(t, t)
        else:
            v = t
        if ab: v = f"|{v}|"
        if neg: v = f"-{v}"
        return v

    def regs_in(tok):
        return re.findall(r"U?R\d+", tok)

    for addr, ins in lines:
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        m = re.match(r"(\S+)\s*(.*)", ins)
        op, rest = m.group(1), m.group(2)
        toks = [t.strip() for t in re.split(r",\s*(?![^\[]*\])", rest)] if rest else []
        base = op.split(".")[0]
        if base in ("STG", "ST"):
            m2 = re.search(r"\[(R\d+)(?:\.64)?(?:\+0x([0-9a-f]+))?\]$", toks[0])
            areg = m2.group(1) if m2 else "?"
            off = int(m2.group(2), 16) if (m2 and m2.group(2)) else 0
            width = 4
            if ".64" in op: width = 8
            if ".128" in op: width = 16
            r0 = int(re.match(r"R(\d+)", toks[1]).group(1)) if re.match(r"R\d+", toks[1]) else None
            for k in range(width // 4):
                r = f"R{r0 + k}" if r0 is not None else toks[1]
                used_by_store.append((addr, f"[{tag.get(areg, areg)}+{off + 4 * k}]", val.get(r, r if r0 is not None else "0")))
            continue
        if not toks:
            continue
        dst = toks[0]
        if not re.fullmatch(r"U?R\d+", dst):
            # predicate destinations etc.
            if base in ("FSETP", "DSETP", "ISETP"):
                pd = dst
                val[pd] = new_node(f"{op}({', '.join(src(t) for t in toks[2:])})") if base != "ISETP" else f"icmp@{addr}"
            continue
        srcs = toks[1:]
        if base in ("LDG", "LD", "LDS", "LDL"):
            m2 = re.search(r"\[(U?R\d+)(?:\.64)?(?:\+0x([0-9a-f]+))?\]$", srcs[-1])
            areg = m2.group(1) if m2 else "?"
            off = int(m2.group(2), 16) if (m2 and m2.group(2)) else 0
            width = 4
            if ".64" in op: width = 8
            if ".128" in op: width = 16
            d0 = int(dst[1:])
            for k in range(width // 4):
                val[f"R{d0 + k}"] = f"{base}[{tag.get(areg, areg)}:{areg}+{off + 4 * k}]"
                tag.pop(f"R{d0 + k}", None)
            continue
        if base in ("LDC", "LDCU", "ULDC"):
            m2 = re.match(r"c\[0x0\]\[0x([0-9a-f]+)\]", srcs[0])
            d0 = int(re.search(r"\d+", dst).group())
            pre = "UR" if dst.startswith("UR") else "R"
            if m2:
                o = int(m2.group(1), 16) - pbase
                n = 2 if ".64" in op else 1
                for k in range(n):
                    val[f"{pre}{d0 + k}"] = f"P@{o + 4 * k}"
                    tag[f"{pre}{d0 + k}"] = f"P@{o}"
            else:
                val[dst] = f"{op}({rest})"
            continue
        is_fp = base in FP_OPS
        if base in ("MOV", "UMOV") or op.startswith("IMAD.MOV"):
            s = srcs[-1] if base != "IMAD" else srcs[-1]
            if op.startswith("IMAD.MOV"):
                s = srcs[2]
            val[dst] = src(s)
            rs = regs_in(s)
            if rs and rs[0] in tag: tag[dst] = tag[rs[0]]
            else: tag.pop(dst, None)
            continue
        if is_fp:
            text = f"{op}({', '.join(src(t) for t in srcs)})"
            n = new_node(text)
            val[dst] = n
            tag.pop(dst, None)
            if base in ("DFMA", "DMUL", "DADD", "DMNMX") or op.startswith("F2F.F64") or "64H" in op or op.startswith("I2F.F64"):
                d0 = int(dst[1:])
                if "64H" in op:
                    val[dst] = n          # high word only
                else:
                    val[f"R{d0 + 1}"] = f"hi({n})"
            continue
        # integer / other: propagate provenance tag, keep opaque value
        t_in = None
        for s in srcs:
            for r in regs_in(s):
                if r in tag and t_in is None: t_in = tag[r]
            m2 = re.match(r"c\[0x0\]\[0x([0-9a-f]+)\]", s.strip())
            if m2 and t_in is None: t_in = f"P@{int(m2.group(1), 16) - pbase}"
        val[dst] = f"i<{op}@{addr}>"
        if t_in: tag[dst] = t_in
        else: tag.pop(dst, None)
        if ".WIDE" in op or ".64" in op:
            d0 = int(re.search(r"\d+", dst).group())
            pre = "UR" if dst.startswith("UR") else "R"
            val[f"{pre}{d0 + 1}"] = f"i<{op}@{addr}.hi>"
            if t_in: tag[f"{pre}{d0 + 1}"] = t_in

    # prune: print only nodes reachable from stores
    printed = set()

    def emit(n):
        if not isinstance(n, str): return
        for tok in re.findall(r"t\d+", n):
            if tok in nodes and tok not in printed:
                printed.add(tok)
                emit(nodes[tok])
                print(f"  {tok} = {nodes[tok]}")

    for addr, where, v in used_by_store:
        emit(v)
        print(f"STORE@{addr} {where} <- {v}")


if __name__ == "__main__":
    main()
