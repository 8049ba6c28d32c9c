#!/usr/bin/env python
"""Pruned SSA listing of the floating-point dataflow behind the stores of one PTX kernel.

Used to pin the *order of FP operations and FMA contraction* nvcc chose (the contraction is decided
before PTX: fma.rn.f32 vs mul.f32/add.f32), so that hand-written kernels can state the same sequence
with explicit __fmaf_rn/__fmul_rn/__fadd_rn intrinsics, and so that two PTX files can be compared
without a GPU.  Control flow is ignored: the most recent textual definition of a register wins,
which is exact for the straight-line per-Gaussian kernels.

usage: ptx_expr.py file.ptx kernel_substring [--dst PARAM_INDEX ...]
  --dst N   only stores whose address derives from kernel parameter N
"""
import re, sys, struct

def kernel_body(text, key):
    m = None
    for mm in re.finditer(r"\.entry\s+(\S+)\(", text):
        if key in mm.group(1):
            m = mm; break
    if m is None:
        raise SystemExit(f"kernel containing {key!r} not found")
    start = text.index("{", m.end())
    depth, i = 0, start
    while True:
        c = text[i]
        if c == "{": depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0: break
        i += 1
    return m.group(1), text[start + 1:i]

FLOATY = re.compile(r"\.(f32|f64)")

def const(r):
    if r.startswith("0f"): return repr(struct.unpack(">f", bytes.fromhex(r[2:]))[0]) + "f"
    if r.startswith("0d"): return repr(struct.unpack(">d", bytes.fromhex(r[2:]))[0]) + "d"
    return r

def main():
    path, key = sys.argv[1], sys.argv[2]
    dsts = [int(sys.argv[i + 1]) for i, a in enumerate(sys.argv) if a == "--dst"]
    name, body = kernel_body(open(path).read(), key)
    defs = {}      # reg -> (op, [args]) ; latest textual definition
    version = {}
    stores = []
    def base_param(r, seen=0):
        """which kernel param an address register derives from, plus constant offset pieces"""
        r = r.strip()
        if seen > 40 or r not in defs: return None
        op, a = defs[r]
        if op == "param": return a[0]
        for x in a:
            p = base_param(x, seen + 1)
            if p is not None: return p
        return None
    lines = body.split("\n")
    prog = []
    for line in lines:
        line = line.strip().rstrip(";")
        if not line or line.startswith("//") or line.startswith(".") or line.endswith(":"): continue
        line = re.sub(r"^@!?%p\d+\s+", "", line)
        m = re.match(r"(\S+)\s+(.*)", line)
        if not m: continue
        op, rest = m.group(1), m.group(2)
        args = [a.strip() for a in re.split(r",\s*(?![^\[]*\])(?![^{]*})", rest)]
        if op.startswith("ld.param"):
            pp = args[1].strip("[]").split("_param_")[-1]
            defs[args[0]] = ("param", [int(pp) if pp.isdigit() else pp])
        elif re.match(r"ld\.(global|shared|const)", op):
            addr = args[-1].strip("[]")
            base, _, off = addr.partition("+")
            p = base_param(base)
            tag = f"{'G' if 'global' in op else 'S'}{p if p is not None else '?'}[{base.strip().lstrip('%')}+{off or 0}]"
            regs = args[0].strip("{}").split(",") if args[0].startswith("{") else [args[0]]
            for k, rr in enumerate(regs):
                defs[rr.strip()] = ("load", [f"{tag}" + (f".{k}" if len(regs) > 1 else "")])
        elif re.match(r"st\.(global|shared)", op):
            addr = args[0].strip("[]")
            base, _, off = addr.partition("+")
            p = base_param(base)
            vals = args[1].strip("{}").split(",") if args[1].startswith("{") else [args[1]]
            for k, v in enumerate(vals):
                stores.append((p, f"{'G' if 'global' in op else 'S'}{p}+{off or 0}" + (f".{k}" if len(vals) > 1 else ""), v.strip(), dict(defs)))
        elif op.startswith("mov") or op.startswith("cvta"):
            defs[args[0]] = ("mov", [args[1]])
        elif op.startswith(("bra", "bar", "ret", "call", "st.", "membar", "red", "atom")):
            pass
        else:
            defs[args[0]] = (op, args[1:])
    names = {}
    active = set()
    emitted = []
    def walk(r, d):
        r = r.strip()
        if r.startswith(("0f", "0d")): return const(r)
        if r not in d: return r
        op, a = d[r]
        key = (r, id(d.get(r)))
        if key in names: return names[key]
        if op == "param": s = f"P{a[0]}"; names[key] = s; return s
        if op == "load": names[key] = a[0]; return a[0]
        if op == "mov": s = walk(a[0], d); names[key] = s; return s
        if key in active: return f"loop<{r}>"
        active.add(key)
        ins = [walk(x, d) for x in a]
        active.discard(key)
        if not FLOATY.search(op) and not op.startswith(("selp", "setp", "cvt")):
            s = f"{op}({', '.join(ins)})"
            if len(s) > 40: s = f"int<{r}>"
            names[key] = s; return s
        n = f"t{len(emitted)}"
        names[key] = n
        emitted.append(f"{n} = {op}({', '.join(ins)})")
        return n
    print("//", name)
    for p, tag, v, d in stores:
        if dsts and p not in dsts: continue
        before = len(emitted)
        s = walk(v, d)
        for e in emitted[before:]: print("  " + e)
        print(f"STORE {tag} <- {s}")

if __name__ == "__main__":
    main()
