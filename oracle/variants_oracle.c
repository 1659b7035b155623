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

/* ---- CPACK: CPACK.cpp:7-101, pattern lengths CPACK.h:127.  The 16-entry FIFO dictionary (zero-initialised,
 * CPACK.h:107-113) persists across lines: strictly sequential. ---------------------------------------------------- */
typedef struct { uint8_t e[16][4]; int head; } orc_cpack_state;

void orc_cpack_init(orc_cpack_state* s) { memset(s, 0, sizeof(*s)); }

/* counts[6]: ZZZZ, XXXX, MMMM, MMXX, ZZZX, MMMX in the order of m_PatternLength (CPACK.h:119-127) */
unsigned orc_cpack_line(orc_cpack_state* s, const uint8_t* line, unsigned L, uint64_t* counts) {
  static const unsigned len[6] = {2, 34, 6, 24, 12, 16};
  unsigned size = 0;
  for (unsigned i = 0; i < L / 4; i++) {
    const uint8_t* w = line + 4 * i;
    int pat = -1;
    if (w[0] == 0 && w[1] == 0 && w[2] == 0) {
      pat = (w[3] == 0) ? 0 : 4;
    } else {
      for (int j = 0; j < 16 && pat < 0; j++) { /* oldest entry first (deque front), CPACK.cpp:41 */
        const uint8_t* d = s->e[(s->head + j) & 15];
        if (w[0] == d[0] && w[1] == d[1]) pat = (w[2] == d[2]) ? ((w[3] == d[3]) ? 2 : 5) : 3;
      }
      if (pat < 0) { /* xxxx: push back, pop front (CPACK.cpp:84-93) */
        pat = 1;
        memcpy(s->e[s->head], w, 4);
        s->head = (s->head + 1) & 15;
      }
    }
    size += len[pat];
    if (counts) counts[pat]++;
  }
  return size;
}

void orc_cpack_run(const uint8_t* lines, uint64_t n, unsigned L, uint32_t* sizes, uint64_t* counts) {
  orc_cpack_state s;
  orc_cpack_init(&s);
  for (uint64_t i = 0; i < n; i++) {
    unsigned v = orc_cpack_line(&s, lines + i * L, L, counts);
    if (sizes) sizes[i] = v;
  }
}

/* ---- SC2: SC2.cpp:270-334 (sampling, trim to 1024 symbols, tree, lookup), heap SC2.cpp:24-126 --------------------- */
#include <stdlib.h>

typedef struct orc_node { int64_t symbol; uint64_t freq; struct orc_node *left, *right; } orc_node;

static void heap_swap(orc_node** a, int i, int j) { orc_node* t = a[i]; a[i] = a[j]; a[j] = t; }
static void heapify(orc_node** a, int size, int index) { /* SC2.cpp:61-84 */
  int m = index, l = 2 * index + 1, r = 2 * index + 2;
  if (l <= size - 1 && a[l]->freq < a[m]->freq) m = l;
  if (r <= size - 1 && a[r]->freq < a[m]->freq) m = r;
  if (m != index) { heap_swap(a, index, m); heapify(a, size, m); }
}
static void leaf_depths(orc_node* nd, int depth, uint32_t* syms, uint8_t* lens, int* k) { /* SC2.cpp:150-162 */
  if (!nd->left && !nd->right) { syms[*k] = (uint32_t)nd->symbol; lens[*k] = (uint8_t)depth; (*k)++; return; }
  leaf_depths(nd->left, depth + 1, syms, lens, k);
  leaf_depths(nd->right, depth + 1, syms, lens, k);
}
static int cmp_u32(const void* a, const void* b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }
typedef struct { uint32_t sym; uint64_t freq; } orc_sf;
static int cmp_freq_sym(const void* a, const void* b) { /* huffman::cmp, SC2.cpp:257-261 */
  const orc_sf *x = a, *y = b;
  if (x->freq == y->freq) return x->sym < y->sym ? -1 : x->sym > y->sym;
  return x->freq < y->freq ? -1 : 1;
}
static int cmp_sym(const void* a, const void* b) { const orc_sf *x = a, *y = b; return x->sym < y->sym ? -1 : x->sym > y->sym; }
typedef struct { uint32_t sym; uint8_t len; } orc_code;
static int cmp_code(const void* a, const void* b) { const orc_code *x = a, *y = b; return x->sym < y->sym ? -1 : x->sym > y->sym; }

/* Builds the code-length table from the words of the first `sampling` lines.  Returns the number of symbols. */
int orc_sc2_table(const uint8_t* lines, uint64_t sampling, unsigned L, uint32_t* out_syms, uint8_t* out_lens) {
  const uint64_t nw = sampling * (L / 4);
  uint32_t* w = (uint32_t*)malloc(nw * 4 + 4);
  memcpy(w, lines, nw * 4);
  qsort(w, nw, 4, cmp_u32);
  orc_sf* sf = (orc_sf*)malloc((nw + 1) * sizeof(orc_sf));
  uint64_t d = 0;
  for (uint64_t i = 0; i < nw;) {
    uint64_t j = i;
    while (j < nw && w[j] == w[i]) j++;
    sf[d].sym = w[i]; sf[d].freq = j - i; d++;
    i = j;
  }
  free(w);
  if (d > 1024) { /* drop the least frequent (ties: smaller symbol first) until 1024 remain, SC2.cpp:294-307 */
    qsort(sf, d, sizeof(orc_sf), cmp_freq_sym);
    memmove(sf, sf + (d - 1024), 1024 * sizeof(orc_sf));
    d = 1024;
    qsort(sf, d, sizeof(orc_sf), cmp_sym); /* std::map order */
  }
  if (d == 0) { free(sf); return 0; }
  orc_node** heap = (orc_node**)malloc(1025 * sizeof(orc_node*));
  orc_node* pool = (orc_node*)calloc(2 * 1024 + 2, sizeof(orc_node));
  int np = 0, size = (int)d;
  for (int i = 0; i < size; i++) { pool[np].symbol = sf[i].sym; pool[np].freq = sf[i].freq; heap[i] = &pool[np++]; }
  for (int i = size / 2 - 1; i >= 0; i--) heapify(heap, size, i); /* buildHeap, SC2.cpp:49-59 */
  while (size > 1) { /* BuildHuffmanTree, SC2.cpp:138-148 */
    orc_node* l = heap[0]; heap_swap(heap, 0, size - 1); size--; heapify(heap, size, 0);
    orc_node* r = heap[0]; heap_swap(heap, 0, size - 1); size--; heapify(heap, size, 0);
    orc_node* nn = &pool[np++];
    nn->symbol = -1; nn->freq = l->freq + r->freq; nn->left = l; nn->right = r;
    heap[size++] = nn; /* AddNode, SC2.cpp:97-111 */
    for (int i = size - 1; i > 0;) {
      int p = (i + 1) / 2 - 1; /* ceil(i / 2) - 1 */
      if (!(heap[p]->freq > heap[i]->freq)) break;
      heap_swap(heap, i, p);
      i = p;
    }
  }
  int k = 0;
  leaf_depths(heap[0], 0, out_syms, out_lens, &k);
  orc_code* codes = (orc_code*)malloc((size_t)k * sizeof(orc_code));
  for (int i = 0; i < k; i++) { codes[i].sym = out_syms[i]; codes[i].len = out_lens[i]; }
  qsort(codes, (size_t)k, sizeof(orc_code), cmp_code);
  for (int i = 0; i < k; i++) { out_syms[i] = codes[i].sym; out_lens[i] = codes[i].len; }
  free(codes); free(pool); free(heap); free(sf);
  return k;
}

/* sizes per line.  `sampling` = number of sampling lines (main.cpp:108-114 computes it from the loader's row count). */
void orc_sc2_run(const uint8_t* lines, uint64_t n, unsigned L, uint64_t sampling, uint32_t* sizes) {
  const unsigned W = L / 4;
  uint32_t syms[1024];
  uint8_t lens[1024];
  int k = -1;
  for (uint64_t i = 0; i < n; i++) {
    unsigned s = 0;
    if (i < sampling) {
      s = 33 * W;
    } else {
      if (k < 0) k = orc_sc2_table(lines, sampling, L, syms, lens);
      for (unsigned j = 0; j < W; j++) {
        uint32_t v;
        memcpy(&v, lines + i * L + 4 * j, 4);
        int lo = 0, hi = k - 1, hit = -1;
        while (lo <= hi) { int mid = (lo + hi) / 2; if (syms[mid] == v) { hit = mid; break; } if (syms[mid] < v) lo = mid + 1; else hi = mid - 1; }
        s += hit >= 0 ? lens[hit] : 33;
      }
    }
    sizes[i] = s;
  }
}

/* ---- PATTERN (analysis tool): Pattern.cpp:6-75 (per line), 109-199 (checkPattern == checkBDI), 201-320 (countPattern),
 * Pattern.h:62-125 (counters, byte histograms), LRU.h:17-56 (temporal locality).  isExistedBefore (Pattern.cpp:101-107)
 * never promotes a hit (exist() does not touch the list) and inserts only on a miss, so the cache is a FIFO set of the
 * last `capacity` distinct lines that missed (reference: CACHESIZE = 2^24 - 1, LRU.h:6).
 * stats layout (uint64): [0] Z  [1] R  [2] T  [3] U  [4] Total  [5..10] Implicit  [11..16] Explicit
 *                        [17..272] SymbolCounts  [273..528] SymbolCountsExceptAllZerosAllWordSame ------------------- */
#include <stdlib.h>
#define ORC_PATTERN_WORDS 529

static uint64_t pat_hash(const uint8_t* line, unsigned L) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned i = 0; i < L; i++) { h ^= line[i]; h *= 1099511628211ull; }
  return h ^ (h >> 29);
}

void orc_pattern_run(const uint8_t* lines, uint64_t n, unsigned L, uint64_t capacity, uint32_t* sizes, uint64_t* st) {
  static const unsigned B[6] = {8, 8, 8, 4, 4, 2}, D[6] = {1, 2, 4, 1, 2, 1};
  memset(st, 0, ORC_PATTERN_WORDS * sizeof(uint64_t));
  /* FIFO set: open-addressing table of line indices (+1), tombstones on eviction, and the insertion order */
  uint64_t slots = 16;
  while (slots < 4 * (n < capacity ? n : capacity) + 16) slots <<= 1;
  uint64_t* table = (uint64_t*)calloc(slots, sizeof(uint64_t)); /* 0 empty, ~0 tombstone, else index + 1 */
  uint64_t* fifo = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
  uint64_t head = 0, tail = 0, live = 0, used = 0;
  for (uint64_t i = 0; i < n; i++) {
    const uint8_t* line = lines + i * L;
    int zero = 1, rep = 1;
    for (unsigned k = 0; k < L; k++) if (line[k]) { zero = 0; break; }
    for (unsigned k = 4; k < L; k++) if (line[k] != line[k % 4]) { rep = 0; break; } /* isRepeated(.., 4) */
    if (zero) st[0] += L;
    if (rep) st[1] += L;
    { /* isExistedBefore */
      uint64_t h = pat_hash(line, L), pos = h & (slots - 1), first_tomb = ~0ull;
      int hit = 0;
      while (table[pos] != 0) {
        if (table[pos] == ~0ull) { if (first_tomb == ~0ull) first_tomb = pos; }
        else if (memcmp(lines + (table[pos] - 1) * L, line, L) == 0) { hit = 1; break; }
        pos = (pos + 1) & (slots - 1);
      }
      if (hit) {
        st[2] += L;
      } else {
        if (first_tomb != ~0ull) pos = first_tomb; else used++;
        table[pos] = i + 1;
        fifo[tail++] = i;
        live++;
        while (live > capacity) { /* LRUCache::clean: drop the oldest inserted line */
          const uint64_t victim = fifo[head++];
          uint64_t p = pat_hash(lines + victim * L, L) & (slots - 1);
          while (table[p] != victim + 1) p = (p + 1) & (slots - 1);
          table[p] = ~0ull;
          live--;
        }
        if (used * 2 > slots) { /* rebuild without tombstones */
          memset(table, 0, slots * sizeof(uint64_t));
          used = 0;
          for (uint64_t q = head; q < tail; q++) {
            uint64_t p = pat_hash(lines + fifo[q] * L, L) & (slots - 1);
            while (table[p] != 0) p = (p + 1) & (slots - 1);
            table[p] = fifo[q] + 1;
            used++;
          }
        }
      }
    }
    unsigned best = 8u * L, cur;
    int sel = 9;
    for (int k = 0; k < 6; k++) {
      cur = bdi_check(line, L, B[k], D[k]);
      if (best > cur) { sel = k; best = cur; }
    }
    if (best == 8u * L) sel = 9;
    if (sel == 9) {
      st[3] += L;
    } else { /* countPattern: immediates are implicit bytes, everything else explicit */
      const unsigned bs = B[sel], ds = D[sel], m = L / bs;
      const uint64_t limit = ds == 1 ? 0xffull : ds == 2 ? 0xffffull : 0xffffffffull;
      for (unsigned q = 0; q < m; q++) {
        uint64_t t = 0;
        for (int j = (int)bs - 1; j >= 0; j--) t = (t << 8) | line[q * bs + (unsigned)j];
        if (bdi_reduce_sign(t) <= limit) st[5 + sel] += bs; else st[11 + sel] += bs;
      }
    }
    for (unsigned k = 0; k < L; k++) st[17 + line[k]]++;
    if (!(zero || rep)) for (unsigned k = 0; k < L; k++) st[273 + line[k]]++;
    st[4] += L;
    if (sizes) sizes[i] = best + 4u;
  }
  free(table);
  free(fifo);
}
