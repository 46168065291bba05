#!/usr/bin/env python3
"""TEST-INFRASTRUCTURE TOOL (not product, not run by tests).

Reads the PTX that nvcc emits for the UNMODIFIED reference (oracle/_ref/ref_chunk.ptx, built by
`make -C oracle ptx`) and prints, for a chosen register or for every global store of a kernel,
the fp32 expression tree that produces it, with the inlined GLM simplex bodies collapsed to
S2(x,y) / S3(x,y,z) nodes.  It exists to answer one question while restating the reference's
arithmetic: "which multiplies did the compiler fuse into FMAs at this call site?" - the answer
is context dependent ("FMA contraction"), so it is read from the compiled reference
rather than guessed.

Notation: fma(a,b,c) is an explicit fma.rn.f32 in the PTX (fused by NVVM); a plain `mul`
whose single consumer is a plain add/sub is printed as MULF(...) - ptxas fuses exactly those
into FFMA as well (checked against the SASS); MUL(...) stays a separate FMUL.  Registers with more than one definition (values merged across branches) are
printed as PHI%reg.

usage: ptx_expr.py file.ptx kernel_substring [--reg %f123 ...] [--depth N] [--stores]
"""
import re
import struct
import sys
from collections import defaultdict

FCONST = re.compile(r"^0[fF]([0-9A-Fa-f]{8})$")


def fconst(tok):
    m = FCONST.match(tok)
    if not m:
        return None
    v = struct.unpack("<f", struct.pack("<I", int(m.group(1), 16)))[0]
    return v


class Kernel:
    def __init__(self, lines):
        self.defs = defaultdict(list)  # reg -> [(op, [srcs], lineno)]
        self.uses = defaultdict(int)
        self.consumers = defaultdict(list)
        self.stores = []
        for ln, raw in lines:
            s = raw.strip()
            if not s or s.startswith("//") or s.startswith(".") or s.endswith(":") or s.startswith("{") or s.startswith("}"):
                continue
            if s.startswith("@"):
                s = s.split(None, 1)[1] if " " in s or "\t" in s else s
            s = s.rstrip(";")
            parts = s.split(None, 1)
            if len(parts) < 2:
                continue
            op, rest = parts
            args = [a.strip() for a in rest.replace("{", "").replace("}", "").split(",")]
            if op.startswith("st."):
                self.stores.append((op, args, ln))
                for a in args[1:]:
                    self.uses[a] += 1
                continue
            if op.startswith("bra") or op.startswith("ret") or op.startswith("bar") or op.startswith("call"):
                continue
            dst, srcs = args[0], args[1:]
            self.defs[dst].append((op, srcs, ln))
            for a in srcs:
                a = a.strip("[]")
                a = a.split("+")[0]
                self.uses[a] += 1
                self.consumers[a].append(op)

    # ---- pattern helpers
    def single(self, reg):
        d = self.defs.get(reg)
        if d and len(d) == 1:
            return d[0]
        return None

    def is_op(self, reg, prefix):
        d = self.single(reg)
        return d is not None and d[0].startswith(prefix)

    def match_s2(self, reg):
        """reg = d (pre-scale simplex2 sum). Return (vx, vy) or None."""
        try:
            op, (g2, m2, inner), _ = self.single(reg)
            if not op.startswith("fma"):
                return None
            op, (g0, m0, prod), _ = self.single(inner)
            if not op.startswith("fma"):
                return None
            # g0 = fma(x0y, h0, mul(x0x, a0))
            op, (x0y, h0, mulreg), _ = self.single(g0)
            if not op.startswith("fma"):
                return None
            op, (x0x, a0), _ = self.single(mulreg)
            # x0x = add(sub(vx, ix), t)
            op, (subx, t), _ = self.single(x0x)
            if not op.startswith("add"):
                return None
            op, (vx, ix), _ = self.single(subx)
            if not op.startswith("sub"):
                return None
            op, (suby, t2), _ = self.single(x0y)
            op, (vy, iy), _ = self.single(suby)
            # t must be the unskew dot with 0f3E58658C
            op, tsrc, _ = self.single(t)
            if "0f3E58658C" not in tsrc:
                return None
            return vx, vy
        except Exception:
            return None

    def match_s3(self, reg):
        """reg = add(b, a) pre-scale simplex3 sum. Return (vx,vy,vz) or None."""
        try:
            op, (b, a), _ = self.single(reg)
            if not op.startswith("add"):
                return None
            op, (m1, d1, prod0), _ = self.single(a)
            if not op.startswith("fma"):
                return None
            op, (m0, d0), _ = self.single(prod0)
            # d0 = fma(x0z, p0z, fma(x0x, p0x, mul(x0y, p0y)))
            op, (x0z, p0z, in1), _ = self.single(d0)
            op, (x0x, p0x, in2), _ = self.single(in1)
            op, (x0y, p0y), _ = self.single(in2)
            out = []
            for x0 in (x0x, x0y, x0z):
                op, (sub, t), _ = self.single(x0)
                op2, tsrc, _ = self.single(t)
                if "0f3E2AAAAB" not in tsrc:
                    return None
                op, (v, i), _ = self.single(sub)
                out.append(v)
            return tuple(out)
        except Exception:
            return None

    def match_trig(self, srcs):
        """selp(t, 0-t, p) at the end of libdevice sinf/cosf -> (name, arg register)."""
        try:
            t, nt, p = srcs
            dn = self.single(nt)
            z = dn[1][0]
            dz = self.single(z)
            if dz and dz[0].startswith("mov"):
                z = dz[1][0]
            if not (dn[0].startswith("sub") and fconst(z) == 0.0 and dn[1][1] == t):
                return None
            op, (a, b, c), _ = self.single(t)          # fma(poly, s*r_or_s, r_or_1)
            if not op.startswith("fma"):
                return None
            dc = self.single(c)                          # selp(1.0, r, p) sin | selp(r, 1.0, p) cos
            if not dc[0].startswith("selp"):
                return None
            if fconst(dc[1][0]) == 1.0:
                name, r = "SIN", dc[1][1]
            elif fconst(dc[1][1]) == 1.0:
                name, r = "COS", dc[1][0]
            else:
                return None
            cands = []
            for dop, dsrcs, _ in self.defs.get(r, []):
                cands.append((dop, dsrcs))
                if dop.startswith("selp"):
                    for x in dsrcs[:2]:
                        dx = self.single(x)
                        if dx:
                            cands.append((dx[0], dx[1]))
            for dop, dsrcs in cands:
                if dop.startswith("fma") and dsrcs[1] == "0fA7C234C5":
                    d2 = self.single(dsrcs[2])
                    d1 = self.single(d2[1][2])
                    return name, d1[1][2]
            return None
        except Exception:
            return None

    alias = {}
    tcount = 0
    memo = None
    tdefs = None

    def match_fbm(self, reg):
        """reg = fma(A, amp, C) chain of octaves; returns (n, x0, y0[, z0]) or None."""
        chain = []
        cur = reg
        while True:
            d = self.single(cur)
            if d is None or not d[0].startswith("fma"):
                return None
            a, amp, c = d[1]
            ampv = fconst(amp)
            if ampv is None:
                return None
            da = self.single(a)
            if da is None or not da[0].startswith("mul"):
                return None
            args = None
            for k in (0, 1):
                if da[1][1 - k] == "0f43020000":
                    args = self.match_s2(da[1][k])
                elif da[1][1 - k] == "0f42280000":
                    args = self.match_s3(da[1][k])
            if not args:
                return None
            chain.append((ampv, args))
            if fconst(c) == 0.0:
                break
            cur = c
        chain.reverse()
        amp = 0.5
        for i, (a, args) in enumerate(chain):
            if a != amp:
                return None
            amp *= 0.5
            if i > 0:
                for prev, now in zip(chain[i - 1][1], args):
                    dn = self.single(now)
                    if dn is None or not dn[0].startswith("add") or dn[1][0] != prev or dn[1][1] != prev:
                        return None
        return (len(chain),) + tuple(chain[0][1])

    def expr(self, reg, depth, seen=None):
        c = fconst(reg)
        if c is not None:
            return repr(c)
        if reg in self.alias:
            return self.alias[reg]
        if not reg.startswith("%"):
            return reg
        if self.memo is None:
            self.memo = {}
            self.tdefs = []
        if reg in self.memo:
            return self.memo[reg]
        r = self._expr(reg, depth)
        if len(r) > 100 and self.uses[reg] > 1:
            Kernel.tcount += 1
            name = "t%d" % Kernel.tcount
            self.tdefs.append((name, reg, r))
            r = name
        self.memo[reg] = r
        return r

    def _expr(self, reg, depth):
        if depth <= 0:
            return reg
        ds = self.defs.get(reg)
        if not ds:
            return reg
        if len(ds) > 1:
            return "PHI" + reg
        op, srcs, ln = ds[0]
        E = lambda r: self.expr(r, depth - 1)
        if op.startswith("fma"):
            fb = self.match_fbm(reg)
            if fb:
                return "FBM%d(%s)" % (fb[0], ", ".join(E(x) for x in fb[1:]))
        base = op.split(".")[0]
        if base == "mul" and op.endswith("f32"):
            # simplex scale?
            for k in (0, 1):
                other = srcs[1 - k]
                if other == "0f43020000":
                    m = self.match_s2(srcs[k])
                    if m:
                        return "(130*S2(%s, %s))" % (E(m[0]), E(m[1]))
                if other == "0f42280000":
                    m = self.match_s3(srcs[k])
                    if m:
                        return "(42*S3(%s, %s, %s))" % tuple(E(x) for x in m)
            cons = self.consumers.get(reg, [])
            fusable = (len(cons) == 1 and ".rn" not in op and cons[0] in ("add.f32", "sub.f32"))
            tag = "MULF" if fusable else ("MULrn" if ".rn" in op else "MUL")
            return "%s(%s, %s)" % (tag, E(srcs[0]), E(srcs[1]))
        if base == "fma":
            for (a, b) in ((0, 1), (1, 0)):
                if srcs[b] == "0f43020000":
                    m = self.match_s2(srcs[a])
                    if m:
                        return "fma(S2(%s, %s), 130, %s)" % (E(m[0]), E(m[1]), E(srcs[2]))
                if srcs[b] == "0f42280000":
                    m = self.match_s3(srcs[a])
                    if m:
                        return "fma(S3(%s, %s, %s), 42, %s)" % (E(m[0]), E(m[1]), E(m[2]), E(srcs[2]))
            return "fma(%s, %s, %s)" % (E(srcs[0]), E(srcs[1]), E(srcs[2]))
        if base in ("add", "sub", "div", "min", "max") and ("f32" in op or "f64" in op):
            sym = {"add": "+", "sub": "-", "div": "/"}.get(base)
            rn = "rn" if ".rn" in op and base != "div" else ""
            if sym:
                return "(%s %s%s %s)" % (E(srcs[0]), sym, rn, E(srcs[1]))
            return "%s(%s, %s)" % (base, E(srcs[0]), E(srcs[1]))
        if base == "cvt":
            kind = op.replace("cvt.", "")
            name = {"rmi.f32.f32": "floor", "rpi.f32.f32": "ceil", "rzi.f32.f32": "trunc", "rni.f32.f32": "rint"}.get(kind, "cvt." + kind)
            return "%s(%s)" % (name, E(srcs[0]))
        if base == "selp":
            trig = self.match_trig(srcs)
            if trig:
                return "%s(%s)" % (trig[0], E(trig[1]))
            dp = self.single(srcs[2])
            if dp and dp[0].startswith("setp"):
                cmp_ = dp[0].split(".")[1]
                a, b = dp[1][0], dp[1][1]
                # selp(0, x, x < 0) -> max0(x) ; selp(1, x, x > 1) -> min1(x)
                if cmp_ == "lt" and fconst(b) == 0.0 and fconst(srcs[0]) == 0.0 and srcs[1] == a:
                    return "max0(%s)" % E(a)
                if cmp_ == "gt" and fconst(b) == 1.0 and fconst(srcs[0]) == 1.0 and srcs[1] == a:
                    return "min1(%s)" % E(a)
            return "sel(%s ? %s : %s)" % (E(srcs[2]), E(srcs[0]), E(srcs[1]))
        if base == "setp":
            return "(%s %s %s)" % (E(srcs[0]), op.split(".")[1], E(srcs[1]))
        if base in ("abs", "neg", "sqrt", "rcp", "mov", "ld", "not"):
            return "%s(%s)" % (op if base in ("ld",) else base, ", ".join(E(s) for s in srcs))
        return "%s(%s)" % (op, ", ".join(E(s) for s in srcs))


def load_kernel(path, name):
    lines = open(path).read().split("\n")
    start = None
    for i, l in enumerate(lines):
        if (".entry" in l or l.startswith(".func") or l.startswith(".visible .func")) and name in l:
            start = i
            break
    if start is None:
        raise SystemExit("kernel not found: " + name)
    depth = 0
    body = []
    began = False
    for i in range(start, len(lines)):
        l = lines[i]
        if l.strip() == "{":
            depth += 1
            began = True
        elif l.strip() == "}":
            depth -= 1
            if began and depth == 0:
                break
        body.append((i + 1, l))
    return Kernel(body)


def main():
    path, name = sys.argv[1], sys.argv[2]
    depth = 12
    regs = []
    aliases = []
    do_stores = False
    a = sys.argv[3:]
    i = 0
    while i < len(a):
        if a[i] == "--depth":
            depth = int(a[i + 1]); i += 2
        elif a[i] == "--reg":
            regs.append(a[i + 1]); i += 2
        elif a[i] == "--stores":
            do_stores = True; i += 1
        elif a[i] == "--alias":
            aliases.append(a[i + 1]); i += 2
        else:
            i += 1
    k = load_kernel(path, name)
    for al in aliases:
        r, n = al.split("=")
        k.alias[r] = n
    def flush():
        for name, reg, r in k.tdefs:
            print("   %s [%s] = %s" % (name, reg, r))
        del k.tdefs[:]
    for r in regs:
        e = k.expr(r, depth)
        flush()
        print(r, "=", e)
    if do_stores:
        for op, args, ln in k.stores:
            e = k.expr(args[1], depth)
            flush()
            print("line", ln, op, args[0], "<-", e)


if __name__ == "__main__":
    main()
