#!/usr/bin/env python3
"""Config compiler for the specialised (thread-per-block) MPC kernel.

Reads a reference-format MPC config (the JSON VPC::parseConfig reads, VPC.cpp:72-330) and writes
cal_22-mpc_b200/csrc/spec/spec_<NAME>.cu: a straight-line schedule over the hand-written primitives of
csrc/mpc_spec.cuh with every table folded into immediates --

  * predictor byte gathers (PredictorModule.cpp:37-173) and the root-first residue order
    (ResidueModule.cpp:26-39) become PRMTs with constant selectors or plain register renaming,
  * DiffTable bytes / WeightTable shift distances become immediate operands,
  * the scan permutation (ScanModule.cpp:14-20) is recognised as column-major (a scan row = two residue
    bytes) or plane-major (a scan row = one bit plane of 16 bytes); column-major modules are scored lazily,
    row by row, and stop at their first non-zero row (VPC.cpp:378-387).

A config is eligible when lineSize is 128 and every PredComp scan table is column-major or plane-major;
anything else keeps using the generic warp-per-block kernel.  The emitted file also carries the flattened
config so that the library can check at run time that a context's config is the one compiled in.

usage: gen_spec.py configs/F4.json [...]      (writes one .cu per config + spec/spec_list.inc)
"""
import json
import math
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUTDIR = os.path.join(ROOT, "cal_22-mpc_b200", "csrc", "spec")
L = 128
W = 32


class NotEligible(Exception):
    pass


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def weight_shift(w):
    w = f32(w)
    if not (w > 0.0) or math.isinf(w):
        raise NotEligible("WeightTable entry without a defined shift")
    s = int(math.log2(w))  # (int)log2f(w), PredictorModule.cpp:31 -- truncation toward zero
    # log2f in float: guard the representable powers of two against double rounding
    lf = f32(math.log2(w))
    s = int(lf)
    return max(-8, min(8, s))


# ---------------------------------------------------------------------------------------------------------------
# byte-gather planning
# ---------------------------------------------------------------------------------------------------------------
def gather_expr(arr, srcs):
    """Expression for a 32-bit word whose byte q is byte srcs[q] (a byte index into the 128-byte array `arr`,
    which is available as 32 words arr[0..31]) or zero when srcs[q] is None."""
    if all(s is None for s in srcs):
        return "0u"
    words = []
    for s in srcs:
        if s is not None and s // 4 not in words:
            words.append(s // 4)
    mask = sum(0xFF << (8 * q) for q, s in enumerate(srcs) if s is not None)

    def w(i):
        return f"{arr}[{i}]"

    def finish(e):
        return e if mask == 0xFFFFFFFF else f"({e} & 0x{mask:08x}u)"

    if len(words) == 1 and all(s is None or s % 4 == q for q, s in enumerate(srcs)):
        return finish(w(words[0]))
    if len(words) <= 2:
        a = words[0]
        b = words[1] if len(words) == 2 else words[0]
        sel = 0
        for q, s in enumerate(srcs):
            if s is None:
                nib = 0
            else:
                nib = (s % 4) + (0 if s // 4 == a else 4)
            sel |= nib << (4 * q)
        return finish(f"prmt({w(a)}, {w(b)}, 0x{sel:04x}u)")
    # 3 or 4 source words: two partial gathers merged by a third PRMT
    first, second = words[:2], words[2:]

    def partial(ws):
        a = ws[0]
        b = ws[1] if len(ws) == 2 else ws[0]
        sel = 0
        for q, s in enumerate(srcs):
            nib = 0
            if s is not None and s // 4 in ws:
                nib = (s % 4) + (0 if s // 4 == a else 4)
            sel |= nib << (4 * q)
        if len(ws) == 1 and sel == 0x3210:
            return w(a)
        return f"prmt({w(a)}, {w(b)}, 0x{sel:04x}u)"

    sel = 0
    for q, s in enumerate(srcs):
        nib = q if (s is None or s // 4 in first) else 4 + q
        sel |= nib << (4 * q)
    return finish(f"prmt({partial(first)}, {partial(second)}, 0x{sel:04x}u)")


# ---------------------------------------------------------------------------------------------------------------
# module model
# ---------------------------------------------------------------------------------------------------------------
class Module:
    def __init__(self, idx, spec):
        self.idx = idx
        sub = spec["submodules"]
        ps = sub["ResidueModule"]["PredictorModule"]
        self.pname = ps["name"]
        self.root = int(ps.get("RootIndex", 0))
        self.cxor = bool(sub["XORModule"].get("consecutiveXOR", False))
        if self.pname in ("DiffBasePredictor", "WeightBasePredictor") and int(ps.get("LineSize", 0)) != L:
            raise NotEligible("predictor LineSize != 128")
        if self.pname == "ConsecutiveBasePredictor" and self.root != 0:
            raise NotEligible("Consecutive predictor with root != 0")
        # ---- source tables in residue order (same construction as build_generic_tables, mpc_generic.cu) ----
        tperm = []
        for plane in (3, 2, 1, 0):
            tperm += list(range(plane, L, 4))
        self.xsrc, self.psrc, self.pval = [], [], []
        for j in range(L):
            i = self.root if j == 0 else (j - 1 if j <= self.root else j)
            self.xsrc.append(i)
            if self.pname == "OneBasePredictor":
                ps_, pv = self.root, 0
            elif self.pname == "ConsecutiveBasePredictor":
                ps_, pv = tperm[i - 1 if i > 0 else 0], 0
            elif self.pname == "DiffBasePredictor":
                ps_, pv = int(ps["BaseIndexTable"][i]), int(ps["DiffTable"][i]) & 0xFF
            elif self.pname == "WeightBasePredictor":
                ps_, pv = int(ps["BaseIndexTable"][i]), (weight_shift(ps["WeightTable"][i]) if i != self.root else 0)
            else:
                raise NotEligible("unknown predictor " + self.pname)
            if j == 0:
                ps_, pv = self.root, 0
            if not (0 <= ps_ < L):
                raise NotEligible("base index outside the line")
            self.psrc.append(ps_)
            self.pval.append(pv)
        self.op = {"DiffBasePredictor": "add", "WeightBasePredictor": "shift"}.get(self.pname, "none")
        self.root_pred = tperm[self.root] if self.pname == "ConsecutiveBasePredictor" else self.root
        # ---- scan family ----
        sc = sub["ScanModule"]
        T = int(sc["TableSize"])
        rows, cols = [int(v) for v in sc["Rows"][:T]], [int(v) for v in sc["Cols"][:T]]
        if any(not (0 <= r <= 7) for r in rows) or any(not (0 <= c < L) for c in cols):
            raise NotEligible("scan entry outside the bit-plane array")
        self.family = None
        if T % 8 == 0 and all(rows[i] == i % 8 for i in range(T)) and \
                all(len(set(cols[i:i + 8])) == 1 for i in range(0, T, 8)):
            self.family = "cm"
            self.cols = [cols[i] for i in range(0, T, 8)]  # scan order; rows beyond are zero
        elif T == 8 * L and all(len(set(rows[i:i + L])) == 1 for i in range(0, T, L)) and \
                sorted(rows[i] for i in range(0, T, L)) == list(range(8)) and \
                all(cols[i:i + L] == cols[:L] for i in range(0, T, L)):
            self.family = "pm"
            self.cols = cols[:L]
            self.rho = [rows[i] for i in range(0, T, L)]  # plane scanned by plane group p
        else:
            raise NotEligible(f"module {idx}: scan table is neither column-major nor plane-major")

    # ---- expressions -----------------------------------------------------------------------------------------
    def pred_expr(self, w):
        srcs = self.psrc[4 * w:4 * w + 4]
        base = gather_expr("x", srcs)
        vals = self.pval[4 * w:4 * w + 4]
        if self.op == "add":
            d = sum((v & 0xFF) << (8 * q) for q, v in enumerate(vals))
            return base if d == 0 else f"add_u8x4({base}, 0x{d:08x}u)"
        if self.op == "shift":
            groups = {}
            for q, s in enumerate(vals):
                groups.setdefault(s, []).append(q)
            if list(groups.keys()) == [0]:
                return base
            terms = []
            for s, qs in sorted(groups.items()):
                if abs(s) >= 8:
                    continue
                if s >= 0:
                    m = sum((((0xFF << s) & 0xFF) << (8 * q)) for q in qs)
                    terms.append(f"((p << {s}) & 0x{m:08x}u)" if s else f"(p & 0x{m:08x}u)")
                else:
                    m = sum(((0xFF >> -s) << (8 * q)) for q in qs)
                    terms.append(f"((p >> {-s}) & 0x{m:08x}u)")
            if not terms:
                return "0u"
            return f"shiftmix({base}, [](uint32_t p) {{ return {' | '.join(terms)}; }})"
        return base

    def residue_stmts(self, w, name, shared_low=False):
        """C++ statements defining `const uint32_t <name>` = residue word w (before the XOR stage).
        shared_low: both operands are plain line words and `al[i]` = x[i] & 0x7f7f7f7f is available -- the low-7-bit
        halves of the SWAR subtract are then shared between the word's role as minuend and as prediction."""
        xe = gather_expr("x", self.xsrc[4 * w:4 * w + 4])
        pe = self.pred_expr(w)
        import re as _re
        mx, mp = _re.fullmatch(r"x\[(\d+)\]", xe), _re.fullmatch(r"x\[(\d+)\]", pe)
        if shared_low and mx and mp:
            a, b = mx.group(1), mp.group(1)
            e = f"sub_u8x4_shared(x[{a}], x[{b}], al[{a}], al[{b}])"
        else:
            e = f"sub_u8x4({xe}, {pe})"
        if w == 0:
            rw, rb = self.root // 4, self.root % 4
            rootbyte = f"(x[{rw}] & 0xffu)" if rb == 0 else f"((x[{rw}] >> {8 * rb}) & 0xffu)"
            e = f"(({e} & 0xffffff00u) | {rootbyte})"
        return f"const uint32_t {name} = {e};"

    def eq_expr(self, w):
        """line word ^ predicted word for residue word w; a zero byte <=> that residue byte is zero."""
        xe = gather_expr("x", self.xsrc[4 * w:4 * w + 4])
        e = f"({xe} ^ {self.pred_expr(w)})"
        if w == 0:  # residue position 0 is the root byte itself (ResidueModule.cpp:26-27)
            rw, rb = self.root // 4, self.root % 4
            rootbyte = f"(x[{rw}] & 0xffu)" if rb == 0 else f"((x[{rw}] >> {8 * rb}) & 0xffu)"
            e = f"(({e} & 0xffffff00u) | {rootbyte})"
        return e

    def g_from_r(self, w, r):
        if self.cxor:
            return f"xc({r}, 0x{0x7f7f7f00 if w == 0 else 0x7f7f7f7f:08x}u)"
        return f"xf({r}, 0x{0x01010100 if w == 0 else 0x01010101:08x}u)"


# ---------------------------------------------------------------------------------------------------------------
# emission
# ---------------------------------------------------------------------------------------------------------------
def emit_full(m, out, lut_xor=0):
    """full_<m>: all 32 residue words -> residue sums, canonical row layout c[32].
    lut_xor != 0 (column-major modules only): the row-cost table is indexed with the residue bytes BEFORE the XOR
    stage -- the stage is a per-byte bijection (Gray code / conditional complement), so it is folded into the table
    and disappears from the kernel; only the root byte, which the XOR stage skips (XORModule.cpp:12), is run
    through the inverse map so that the table maps it back to itself."""
    out.append(f"__device__ __forceinline__ void full_{m.idx}(const uint32_t (&x)[32], uint32_t (&c)[32], uint32_t& sa, uint32_t& sq) {{")
    out.append("  uint32_t g[32];")
    out.append("  uint32_t sa0 = 0, sa1 = 0, sa2 = 0, sa3 = 0, sq0 = 0, sq1 = 0, sq2 = 0, sq3 = 0;  // four short chains instead of one long one")
    out.append("  uint32_t al[32];")
    out.append("#pragma unroll")
    out.append("  for (int i = 0; i < 32; i++) al[i] = x[i] & 0x7f7f7f7fu;  // only the words a plain-copy predictor uses survive")
    for w in range(W):
        out.append("  { " + m.residue_stmts(w, "r", shared_low=True))
        if w == 0:
            # MAE/MSE run over all line positions (ResidueModule.cpp:43-73): the residue line holds the root byte
            # itself at position 0, the statistics hold line[root] - predicted[root] instead
            rw, rb = m.root // 4, m.root % 4
            pw, pb = m.root_pred // 4, m.root_pred % 4
            if m.root_pred == m.root:
                out.append("    const uint32_t rs = r & 0xffffff00u;")
            else:
                out.append(f"    const uint32_t rs = (r & 0xffffff00u) | (((x[{rw}] >> {8 * rb}) - (x[{pw}] >> {8 * pb})) & 0xffu);")
            out.append("    sa0 = mpcdev::sum_u8x4_acc(rs, sa0); sq0 = __dp4a(rs, rs, sq0);")
        else:
            out.append(f"    sa{w % 4} = mpcdev::sum_u8x4_acc(r, sa{w % 4}); sq{w % 4} = __dp4a(r, r, sq{w % 4});")
        if lut_xor and m.family == "cm":
            if w == 0:
                inv = "inv_gray8" if lut_xor == 1 else "inv_first8"
                out.append(f"    g[0] = (r & 0xffffff00u) | {inv}(r & 0xffu); }}")
            else:
                out.append(f"    g[{w}] = r; }}")
        else:
            out.append(f"    g[{w}] = {m.g_from_r(w, 'r')}; }}")
    out.append("  sa = (sa0 + sa1) + (sa2 + sa3); sq = (sq0 + sq1) + (sq2 + sq3);")
    if m.family == "cm":
        cols = m.cols + [None] * (L - len(m.cols))
        for j in range(W):
            srcs = [cols[4 * j + 1], cols[4 * j], cols[4 * j + 3], cols[4 * j + 2]]
            out.append(f"  c[{j}] = {gather_expr('g', srcs)};")
    else:
        for h in range(2):
            for k in range(16):
                srcs = [m.cols[16 * (4 * h + q) + k] for q in range(4)]
                out.append(f"  c[{16 * h + k}] = {gather_expr('g', srcs)};")
    out.append("}")


def emit_score_cm(m, out):
    """Leading zero rows of a column-major module.  A scan row is two bytes of the XOR-ed residue line, and such
    a byte is zero exactly when the line byte equals its prediction (both XOR variants map 0 -> 0 only), so the
    score needs no subtraction: e<w> = line word ^ predicted word, tested under the byte masks of the row."""
    out.append(f"__device__ __forceinline__ uint32_t score_{m.idx}(const uint32_t (&x)[32]) {{")
    have = set()
    cols = m.cols + [None] * (L - len(m.cols))
    tests = []
    for k in range(64):
        need = {}
        for cidx in (cols[2 * k], cols[2 * k + 1]):
            if cidx is None:
                continue
            need.setdefault(cidx // 4, 0)
            need[cidx // 4] |= 0xFF << (8 * (cidx % 4))
        tests.append(need)
    k = 0
    while k < 64:
        group = [kk for kk in (k, k + 1) if kk < 64 and tests[kk]]
        for kk in group:
            for w in sorted(tests[kk]):
                if w not in have:
                    out.append(f"  const uint32_t e{w} = {m.eq_expr(w)};")
                    have.add(w)
        exprs = []
        for kk in group:
            terms = {}
            for w, mask in sorted(tests[kk].items()):
                terms.setdefault(mask, []).append(f"e{w}")
            parts = [f"(({' | '.join(ws)}) & 0x{mask:08x}u)" if mask != 0xFFFFFFFF else f"({' | '.join(ws)})"
                     for mask, ws in terms.items()]
            exprs.append((kk, " | ".join(parts)))
        if len(exprs) == 2:
            (k0, e0), (k1, e1) = exprs
            out.append(f"  {{ const uint32_t t0 = {e0}, t1 = {e1}; if ((t0 | t1) != 0u) return t0 ? {k0}u : {k1}u; }}")
        elif len(exprs) == 1:
            out.append(f"  if (({exprs[0][1]}) != 0u) return {exprs[0][0]}u;")
        k += 2
    out.append("  return 64u;")
    out.append("}")


def emit_score_pm(m, out):
    out.append(f"__device__ __forceinline__ uint32_t score_{m.idx}(const uint32_t (&x)[32]) {{")
    out.append("  uint32_t g[32];")
    early = (m.rho[0] == 0)  # plane 0 (bit 7) is untouched by the XOR stage: test it on the residues alone
    chunks = []
    for j in range(8):
        need = {}
        for cidx in m.cols[16 * j:16 * j + 16]:
            need.setdefault(cidx // 4, 0)
            need[cidx // 4] |= 0xFF << (8 * (cidx % 4))
        chunks.append(need)

    def chunk_or(j, arr):
        terms = [f"({arr}[{w}] & 0x{mask:08x}u)" if mask != 0xFFFFFFFF else f"{arr}[{w}]" for w, mask in sorted(chunks[j].items())]
        return " | ".join(terms)

    for w in range(W):
        out.append("  { " + m.residue_stmts(w, "r") + f" g[{w}] = r; }}")
    if early:
        out.append("  {")
        out.append("    uint32_t msb = 0;")
        out.append("#pragma unroll")
        out.append("    for (int i = 0; i < 32; i++) msb |= g[i];")
        out.append("    if (msb & 0x80808080u) {  // some row of the first plane group is non-zero: z < 8")
        for j in range(8):
            out.append(f"      if (({chunk_or(j, 'g')}) & 0x80808080u) return {j}u;")
        out.append("    }")
        out.append("  }")
    if m.rho != list(range(8)):
        for w in range(W):
            out.append(f"  g[{w}] = {m.g_from_r(w, f'g[{w}]')};")
    else:
        out.append("  // planes are scanned MSB first: while every higher plane is zero, plane b of the XOR-ed residue equals")
        out.append("  // plane b of the residue itself (XORModule.cpp:9-20), so the leading-zero count needs no XOR stage")
    out.append("  uint32_t f[8];")
    for j in range(8):
        out.append(f"  {{ uint32_t o = {chunk_or(j, 'g')}; o |= o >> 16; o |= o >> 8; f[{j}] = {plane_order_expr('o', m.rho)}; }}")
    out.append("  return pm_leading_zero_rows(f);")
    out.append("}")


def plane_order_expr(v, rho):
    """8-bit value whose bit 7-p is bit 7-rho[p] of v (planes reordered into scan order)."""
    if rho == list(range(8)):
        return f"({v} & 0xffu)"
    terms = []
    for p, r in enumerate(rho):
        src, dst = 7 - r, 7 - p
        if src >= dst:
            terms.append(f"((({v}) >> {src - dst}) & 0x{1 << dst:02x}u)")
        else:
            terms.append(f"((({v}) << {dst - src}) & 0x{1 << dst:02x}u)")
    return "(" + " | ".join(terms) + ")"


def pm_selectors(rho):
    # after transpose8x8 byte c holds plane 7-c; output byte p must hold plane rho[p]
    sel0 = sum((7 - rho[p]) << (4 * p) for p in range(4))
    sel1 = sum((7 - rho[4 + p]) << (4 * p) for p in range(4))
    return sel0, sel1


def pod_initializer(cfg, mods, n, has_ws, first, enc):
    """C++ aggregate initialiser of the mpc_config_pod the library must see at run time."""
    def arr(vals, width):
        vals = list(vals) + [0] * (width - len(vals))
        return "{" + ",".join(str(v) for v in vals) + "}"

    mod_inits = []
    by_idx = {m.idx: m for m in mods}
    for i in range(16):
        if i >= n:
            mod_inits.append("{}")
            continue
        spec = cfg["modules"][str(i)]
        if spec["name"] == "AllZero":
            mod_inits.append("{MPC_MOD_ALLZERO}")
        elif spec["name"] in ("AllWordSame", "ByteplaneAllSame"):
            mod_inits.append("{MPC_MOD_ALLWORDSAME}")
        else:
            m = by_idx[i]
            sub = spec["submodules"]
            ps = sub["ResidueModule"]["PredictorModule"]
            sc = sub["ScanModule"]
            pid = {"OneBasePredictor": 0, "ConsecutiveBasePredictor": 1, "DiffBasePredictor": 2, "WeightBasePredictor": 3}[m.pname]
            base = [int(v) for v in ps.get("BaseIndexTable", [0] * L)] if pid >= 2 else [0] * L
            base = [b if 0 <= b < L else 0 for b in base]
            diff = [int(v) & 0xFF for v in ps.get("DiffTable", [0] * L)] if pid == 2 else [0] * L
            shift = [(weight_shift(v) if j != m.root else 0) for j, v in enumerate(ps.get("WeightTable", [1.0] * L))] if pid == 3 else [0] * L
            T = int(sc["TableSize"])
            mod_inits.append("{MPC_MOD_PREDCOMP, %d, %d, %d, %d, %s, %s, %s, %s, %s}" % (
                pid, m.root, int(m.cxor), T, arr(base, L), arr(diff, L), arr(shift, L),
                arr([int(v) for v in sc["Rows"][:T]], 8 * L), arr([int(v) for v in sc["Cols"][:T]], 8 * L)))
    return "{%d, %d, %d, %d, %s, {%s}}" % (L, n, has_ws, first, arr(enc, 17), ",\n   ".join(mod_inits))


def generate(cfg_path):
    name = os.path.splitext(os.path.basename(cfg_path))[0]
    with open(cfg_path) as f:
        cfg = json.load(f)
    ov = cfg["overview"]
    n = int(ov["num_modules"])
    if int(ov["lineSize"]) != L:
        raise NotEligible("lineSize != 128")
    names = [cfg["modules"][str(i)]["name"] for i in range(n)]
    has_ws = 0
    for nm in names:
        if nm == "AllZero":
            has_ws = 0
        elif nm in ("AllWordSame", "ByteplaneAllSame"):
            has_ws = 1
    first = 2 if has_ws else 1
    if names[0] != "AllZero" or (has_ws and names[1] not in ("AllWordSame", "ByteplaneAllSame")) or \
            any(nm != "PredComp" for nm in names[first:]):
        raise NotEligible("module order the reference cannot run")
    if ov.get("encoding_bits") is None:
        enc = [int(math.ceil(f32(math.log2(f32(n + 1)))))] * (n + 1)
    else:
        enc = [int(v) for v in ov["encoding_bits"][:n + 1]]
    mods = [Module(i, cfg["modules"][str(i)]) for i in range(first, n)]

    out = []
    out.append(f"// AUTO-GENERATED by tools/gen_spec.py from configs/{name}.json -- do not edit.")
    out.append("// Straight-line schedule of the MPC per-block path for this config over the primitives of mpc_spec.cuh.")
    out.append('#include "../mpc_spec.cuh"')
    out.append('#include "../mpc_spec.h"')
    out.append("")
    out.append("namespace mpc {")
    out.append(f"namespace spec_{name} {{")
    out.append("using namespace mpc::spec;")
    out.append("template <class F> __device__ __forceinline__ uint32_t shiftmix(uint32_t p, F f) { return f(p); }")
    out.append("")
    cm_mods = [m for m in mods if m.family == "cm"]
    lut_on = bool(cm_mods) and os.environ.get("MPC_SPEC_LUT", "1") != "0"
    lut_xor = 0
    if lut_on and os.environ.get("MPC_SPEC_LUTXOR", "1") != "0":
        if all(m.cxor for m in cm_mods):
            lut_xor = 1
        elif not any(m.cxor for m in cm_mods):
            lut_xor = 2
    for m in mods:
        out.append(f"// ---- module {m.idx}: {m.pname}, root {m.root}, {'consecutive' if m.cxor else 'first-plane'} XOR, "
                   f"{'column' if m.family == 'cm' else 'plane'}-major scan ----")
        if m.family == "cm":
            emit_score_cm(m, out)
        else:
            emit_score_pm(m, out)
        emit_full(m, out, lut_xor)
        out.append("")
    out.append("struct Cfg {")
    out.append(f"  static constexpr int kNumModules = {n};")
    out.append(f"  static constexpr int kFirst = {first};")
    out.append(f"  static constexpr bool kHasWordSame = {'true' if has_ws else 'false'};")
    # register budget: plane-major modules keep the line, its residues and the transposed rows live at once
    min_ctas = int(os.environ.get("MPC_SPEC_MIN_CTAS", "0")) or (1 if any(m.family == "pm" for m in mods) else 2)
    has_pm = any(m.family == "pm" for m in mods)
    has_cm = any(m.family == "cm" for m in mods)
    use_lut = has_cm and os.environ.get("MPC_SPEC_LUT", "1") != "0"
    # column-major only: 16 warps in ONE CTA per SM (<= 128 registers) so that the 64 KiB row-cost table, the
    # 128 KiB of tile stages and the histogram fit the 227 KiB of shared memory
    warps = 16 if (use_lut and not has_pm) else 8
    if warps == 16:
        min_ctas = 1
    out.append(f"  static constexpr int kWarps = {warps};")
    out.append(f"  static constexpr bool kUseLut = {'true' if use_lut else 'false'};  // shared-memory row-cost table (column-major modules)")
    skip = os.environ.get("MPC_SPEC_SKIP", "0" if use_lut else "1") != "0"
    out.append(f"  static constexpr bool kSkipZeroGroups = {'true' if skip else 'false'};  // branch around groups of eight zero rows in the encoder")
    out.append(f"  static constexpr int kLutXor = {lut_xor if use_lut else 0};  // 0: table indexed by scan rows; 1 / 2: XOR stage (consecutive / first-plane) folded into the table")
    out.append(f"  static constexpr int kMinCtasPerSm = {min_ctas};  // __launch_bounds__: register budget 65536 / (threads * CTAs)")
    out.append("  __device__ static __forceinline__ uint32_t enc(int k) {  // encoding bits of cluster k-1, VPC.cpp:102-117")
    out.append("    switch (k) {")
    for k, e in enumerate(enc):
        out.append(f"      case {k}: return {e}u;")
    out.append("    }")
    out.append("    return 0u;")
    out.append("  }")
    out.append("  // VPC.cpp:372-395: most leading zero rows wins, ties go to the later module")
    out.append("  __device__ static __forceinline__ void select(const uint32_t (&x)[32], int& best, uint32_t& bestz, unsigned lanes) {")
    out.append("    uint32_t z;")
    for m in mods[:-1]:
        out.append(f"    z = score_{m.idx}(x); if (bestz <= z) {{ best = {m.idx}; bestz = z; }}")
    if mods:
        m = mods[-1]
        out.append("    // the last module wins every tie, so while no earlier module has a zero row it wins unscored")
        out.append(f"    if (bestz == 0u) {{ best = {m.idx}; }} else {{ z = score_{m.idx}(x); if (bestz <= z) {{ best = {m.idx}; bestz = z; }} }}")
        out.append("    __syncwarp(lanes);  // reconverge before the encoder: both sides of the branch share it")
    out.append("  }")
    out.append("  // residue sums (VPC.cpp:417-443) + common encoder (FPCModule.cpp:19-85) of the chosen module")
    out.append("  __device__ static __forceinline__ uint32_t encode(int best, const uint32_t (&x)[32], uint32_t& sa, uint32_t& sq, unsigned lanes, const uint8_t* lut) {")
    out.append("    uint32_t c[32];")
    fams = sorted(set((m.family, pm_selectors(m.rho) if m.family == "pm" else None) for m in mods), key=str)
    out.append("    int fam = 0;")
    out.append("    switch (best) {")
    for m in mods:
        fid = fams.index((m.family, pm_selectors(m.rho) if m.family == "pm" else None))
        out.append(f"      case {m.idx}: full_{m.idx}(x, c, sa, sq); fam = {fid}; break;")
    out.append("      default: break;")
    out.append("    }")
    out.append("    __syncwarp(lanes);  // the row classifier below is shared by all modules: run it once per warp")
    if len(fams) == 1:
        fam, sels = fams[0]
        out.append("    (void)fam;")
        out.append("    return " + ("encode_cm<kUseLut, kSkipZeroGroups>(c, lut);" if fam == "cm" else f"encode_pm<0x{sels[0]:04x}u, 0x{sels[1]:04x}u>(c);"))
    else:
        for fid, (fam, sels) in enumerate(fams):
            call = "encode_cm<kUseLut, kSkipZeroGroups>(c, lut)" if fam == "cm" else f"encode_pm<0x{sels[0]:04x}u, 0x{sels[1]:04x}u>(c)"
            out.append(f"    if (fam == {fid}) return {call};")
        out.append("    return 0u;")
    out.append("  }")
    out.append("};")
    out.append("")
    out.append("static const mpc_config_pod kPod =")
    out.append("  " + pod_initializer(cfg, mods, n, has_ws, first, enc) + ";")
    out.append("")
    out.append("static bool matches(const mpc_config_pod& cfg) { return spec_pod_equal(cfg, kPod); }")
    out.append("static cudaError_t launch(const mpc_config_pod&, const uint8_t* d_lines, uint64_t n_blocks, uint16_t* d_packed,")
    out.append("                          uint64_t* d_stats, const uint8_t* d_row_lut, int sm_count, cudaStream_t stream) {")
    out.append("  return launch_spec<Cfg>(d_lines, n_blocks, d_packed, d_stats, d_row_lut, sm_count, stream);")
    out.append("}")
    out.append(f"}}  // namespace spec_{name}")
    out.append(f'extern const SpecKernel kSpec_{name} = {{"{name}", spec_{name}::matches, spec_{name}::launch, spec_{name}::Cfg::kLutXor}};')
    out.append("}  // namespace mpc")
    os.makedirs(OUTDIR, exist_ok=True)
    path = os.path.join(OUTDIR, f"spec_{name}.cu")
    text = "\n".join(out) + "\n"
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    return name


def main():
    names = []
    for p in sys.argv[1:]:
        try:
            names.append(generate(p))
        except NotEligible as e:
            print(f"gen_spec: {p}: not eligible for the specialised kernel ({e}); the generic kernel serves it", file=sys.stderr)
    lines = ["// AUTO-GENERATED by tools/gen_spec.py: specialised kernels linked into the library."]
    lines += [f"MPC_SPEC({n})" for n in names]
    inc = os.path.join(OUTDIR, "spec_list.inc")
    text = "\n".join(lines) + "\n"
    if not os.path.exists(inc) or open(inc).read() != text:
        with open(inc, "w") as f:
            f.write(text)
    print("gen_spec: generated", " ".join(names))


if __name__ == "__main__":
    main()
