/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the stateless secondary compressor variants (BDI, FPC, BPC) that
 * BASELINE.json config #5 runs over the same dump as MPC.  Plain-C restatement of the reference; never linked
 * into or called from the product library.  Parity status: PINNED -- tests/test_variants.py checks every function
 * per line against the unmodified reference build (oracle/_ref/libmpcref.so) and the known answers of SURVEY.md
 * section 8c.  Paths below are relative to /root/reference/src/compressor/.
 */
#include <stdint.h>
#include <string.h>

/* ---- BDI: BDI.cpp:6-74 (selection), 108-201 (checkBDI), 203-218 (reduceSign) ------------------------------- */

static uint64_t bdi_reduce_sign(uint64_t x) { /* BDI.cpp:203-218 */
  if (x >> 63) {
    for (int i = 62; i >= 0; i--)
      if (((x >> i) & 1) == 0) return x & (0xffffffffffffffffull >> (63 - (i + 1)));
  }
  return x; /* non-negative, or all ones (returned unchanged: never fits a delta) */
}

static unsigned bdi_check(const uint8_t* line, unsigned L, unsigned base_size, unsigned delta_size) {
  const uint64_t limit = delta_size == 1 ? 0xffull : delta_size == 2 ? 0xffffull : 0xffffffffull;
  const unsigned n = L / base_size;
  uint64_t v[128];
  uint8_t imm[128];
  unsigned imm_count = 0;
  for (unsigned i = 0; i < n; i++) { /* little-endian value of the chunk, zero-extended (BDI.cpp:127-153) */
    uint64_t t = 0;
    for (int j = (int)base_size - 1; j >= 0; j--) t = (t << 8) | line[i * base_size + (unsigned)j];
    v[i] = t;
    imm[i] = bdi_reduce_sign(t) <= limit;
    imm_count += imm[i];
  }
  uint64_t base = 0;
  unsigned base_idx = 0;
  for (unsigned i = 0; i < n; i++)
    if (!imm[i]) { base = v[i]; base_idx = i; break; }
  int not_all = 0;
  for (unsigned i = base_idx + 1; i < n; i++)
    if (!imm[i] && bdi_reduce_sign(base - v[i]) > limit) { not_all = 1; break; }
  if (not_all) return n + 8u * (imm_count * delta_size + (n - imm_count) * base_size);
  return n + 8u * (imm_count * delta_size + (base_size + (n - imm_count - 1u) * delta_size)); /* unsigned wrap intended */
}

/* returns size in bits (incl. the 4 encoding bits); *state = BDIState (BDI.h:10-21) */
unsigned orc_bdi_line(const uint8_t* line, unsigned L, int* state) {
  const unsigned raw = 8u * L;
  unsigned best = raw, cur;
  int sel = 8;
  int zero = 1, rep = 1;
  for (unsigned i = 0; i < L; i++) if (line[i]) { zero = 0; break; }
  for (unsigned i = 8; i < L; i++) if (line[i] != line[i % 8]) { rep = 0; break; } /* isRepeated(.., 8), BDI.cpp:86-106 */
  if (zero) { best = 8; sel = 0; }
  else if (rep) { best = 64; sel = 1; }
  else {
    static const unsigned B[6] = {8, 8, 8, 4, 4, 2}, D[6] = {1, 2, 4, 1, 2, 1};
    for (int k = 0; k < 6; k++) {
      cur = bdi_check(line, L, B[k], D[k]);
      if (best > cur) { sel = 2 + k; best = cur; } /* strict: earlier scheme wins ties */
    }
    if (best == raw) sel = 8;
  }
  if (state) *state = sel;
  return best + 4u;
}

/* ---- FPC: FPC.cpp:7-87.  The reference's zero-run scan reads past the line (FPC.cpp:26); here it stops at the line
 * end, which leaves the returned size unchanged and makes the per-word statistics deterministic. ---------------- */
unsigned orc_fpc_line(const uint8_t* line, unsigned L, uint64_t* counts /* [8] per-word prefix counts, may be null */) {
  const unsigned n = L / 4;
  unsigned size = 0, i = 0;
  while (i < n) {
    uint32_t v;
    memcpy(&v, line + 4 * i, 4);
    if (v == 0) {
      size += 6;
      if (counts) counts[0]++;
      i++;
      while (i < n) {
        uint32_t w;
        memcpy(&w, line + 4 * i, 4);
        if (w) break;
        if (counts) counts[0]++;
        i++;
      }
      continue;
    }
    int p;
    if ((v & 0xFFFFFFF8u) == 0 || (v & 0xFFFFFFF8u) == 0xFFFFFFF8u) { size += 7; p = 1; }
    else if ((v & 0xFFFFFF80u) == 0 || (v & 0xFFFFFF80u) == 0xFFFFFF80u) { size += 11; p = 2; }
    else if ((v & 0xFFFF8000u) == 0 || (v & 0xFFFF8000u) == 0xFFFF8000u) { size += 19; p = 3; }
    else if ((v & 0x0000FFFFu) == 0) { size += 19; p = 4; }
    else if ((v & 0xFF80FF80u) == 0 || (v & 0xFF80FF80u) == 0xFF800000u || (v & 0xFF80FF80u) == 0x0000FF80u ||
             (v & 0xFF80FF80u) == 0xFF80FF80u) { size += 19; p = 5; }
    else if ((v & 0xFF) == ((v >> 8) & 0xFF) && (v & 0xFF) == ((v >> 16) & 0xFF) && (v & 0xFF) == (v >> 24)) { size += 11; p = 6; }
    else { size += 35; p = 7; }
    if (counts) counts[p]++;
    i++;
  }
  return size;
}

/* ---- BPC: BPC.cpp:20-87 (planes), 89-101 (encodeFirst: `if (base = 0)` makes it always 7), 103-185 ---------------- */
unsigned orc_bpc_line(const uint8_t* line, unsigned L, uint64_t* pattern_counts /* [7], may be null */, uint64_t* total_words) {
  const unsigned n = L / 4;
  int64_t d[64];
  for (unsigned r = 0; r + 1 < n; r++) { /* words zero-extended to 64 bits (BPC.cpp:41-45 as built with -O3), deltas */
    uint32_t a, b;
    memcpy(&a, line + 4 * r, 4);
    memcpy(&b, line + 4 * (r + 1), 4);
    d[r] = (int64_t)b - (int64_t)a;
  }
  int32_t dbp[33], dbx[33], prev = 0;
  for (int col = 32; col >= 0; col--) {
    int32_t buf = 0;
    for (int r = (int)n - 2; r >= 0; r--) buf = (int32_t)(((uint32_t)buf << 1) | (uint32_t)((d[r] >> col) & 1));
    dbp[col] = buf;
    dbx[col] = (col == 32) ? buf : (buf ^ prev);
    prev = buf;
  }
  unsigned length = 7; /* encodeFirst */
  unsigned run = 0;
  for (int i = 32; i >= 0; i--) {
    if (dbx[i] == 0) { run++; continue; }
    if (run > 0) {
      length += (run == 1) ? 3 : 7;
      if (pattern_counts) pattern_counts[1]++;
      if (total_words) *total_words += run;
    }
    run = 0;
    int pat;
    if (dbp[i] == 0) { length += 5; pat = 2; }
    else if (dbx[i] == 0x7fffffff) { length += 5; pat = 6; }
    else {
      uint32_t u = (uint32_t)dbx[i];
      int ones = __builtin_popcount(u);
      if (ones == 1) { length += 10; pat = 3; }
      else if (ones == 2 && (u & (u >> 1))) { length += 10; pat = 4; }
      else { length += 32; pat = 0; }
    }
    if (pattern_counts) pattern_counts[pat]++;
    if (total_words) *total_words += 1;
  }
  if (run > 0) {
    length += (run == 1) ? 3 : 7;
    if (pattern_counts) pattern_counts[1]++;
    if (total_words) *total_words += run;
  }
  return length;
}

/* alg: 1 BDI, 2 FPC, 3 BPC.  sizes[n]; counts: BDI [9] states, FPC [8] prefixes, BPC [7] patterns + counts[7] = total words */
void orc_variant_run(int alg, const uint8_t* lines, uint64_t n, unsigned L, uint32_t* sizes, uint64_t* counts) {
  for (uint64_t i = 0; i < n; i++) {
    const uint8_t* line = lines + i * L;
    unsigned s = 0;
    if (alg == 1) {
      int st;
      s = orc_bdi_line(line, L, &st);
      if (counts) counts[st]++;
    } else if (alg == 2) {
      s = orc_fpc_line(line, L, counts);
    } else {
      s = orc_bpc_line(line, L, counts, counts ? &counts[7] : 0);
    }
    if (sizes) sizes[i] = s;
  }
}
