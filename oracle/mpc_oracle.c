/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the MPC (VPC) per-block compression path.
 *
 * Plain-C restatement of the reference algorithm (scalable-arch/CAL_22-MPC).  It exists to
 * CHECK the CUDA path; it is never linked into, called from, or shipped with the product
 * library (libmpc_b200.so).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this file per block
 * against the unmodified reference compiled from /root/reference (oracle/_ref/libmpcref.so,
 * recipe oracle/build_ref.sh) and against the committed golden vectors in tests/golden/
 * (generated from that reference build by tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/src/compressor/).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_LINE 256
#define ORC_MAX_MODULES 32

enum { ORC_PRED_ONE = 0, ORC_PRED_CONSEC = 1, ORC_PRED_DIFF = 2, ORC_PRED_WEIGHT = 3 };

typedef struct {
  int predictor, root, consecutive_xor, table_size;
  int base[ORC_MAX_LINE];
  int diff[ORC_MAX_LINE];
  int shift[ORC_MAX_LINE]; /* WeightBase: (int)log2f(weight), PredictorModule.cpp:22-35 */
  int rows[8 * ORC_MAX_LINE];
  int cols[8 * ORC_MAX_LINE];
} orc_predcomp;

typedef struct {
  int line_size;     /* overview.lineSize, VPC.cpp:101 */
  int num_modules;   /* overview.num_modules, VPC.cpp:99 */
  int has_wordsame;  /* module 1 is AllWordSame/ByteplaneAllSame, VPC.cpp:314-319 */
  int first_predcomp;
  int enc_bits[ORC_MAX_MODULES + 1]; /* index = cluster + 1 (cluster -1 = uncompressed), VPC.cpp:102-117 */
  int num_predcomp;
  orc_predcomp pc[ORC_MAX_MODULES];
} orc_config;

orc_config* orc_new(int line_size, int num_modules, int has_wordsame, const int* enc_bits) {
  orc_config* c = (orc_config*)calloc(1, sizeof(orc_config));
  c->line_size = line_size;
  c->num_modules = num_modules;
  c->has_wordsame = has_wordsame;
  c->first_predcomp = has_wordsame ? 2 : 1;
  for (int i = 0; i <= num_modules; i++) c->enc_bits[i] = enc_bits[i];
  return c;
}

/* default encoding bits when the config has no list: ceil(log2f(num_modules + 1)), VPC.cpp:102-108 */
int orc_default_enc_bits(int num_modules) { return (int)ceil(log2f((float)(num_modules + 1))); }

/* WeightBasePredictor constructor, PredictorModule.cpp:22-35 */
int orc_weight_shift(float w) { return (int)log2f(w); }

void orc_add_predcomp(orc_config* c, int predictor, int root, int consecutive_xor, const int* base,
                      const int* diff, const float* weight, int table_size, const int* rows, const int* cols) {
  orc_predcomp* p = &c->pc[c->num_predcomp++];
  p->predictor = predictor;
  p->root = root;
  p->consecutive_xor = consecutive_xor;
  p->table_size = table_size;
  for (int i = 0; i < c->line_size; i++) {
    p->base[i] = base ? base[i] : 0;
    p->diff[i] = diff ? diff[i] : 0;
    p->shift[i] = (weight && i != root) ? orc_weight_shift(weight[i]) : 0;
  }
  for (int i = 0; i < table_size; i++) {
    p->rows[i] = rows[i];
    p->cols[i] = cols[i];
  }
}

void orc_free(orc_config* c) { free(c); }

/* *Predictor::PredictLine, PredictorModule.cpp:37-68 (Weight), 82-110 (Diff), 113-130 (One), 133-173 (Consecutive) */
static void predict(const orc_predcomp* p, const uint8_t* x, int L, uint8_t* out) {
  switch (p->predictor) {
    case ORC_PRED_ONE:
      for (int i = 0; i < L; i++) out[i] = x[p->root];
      break;
    case ORC_PRED_CONSEC: {
      uint8_t t[ORC_MAX_LINE];
      int idx = 0;
      for (int plane = 3; plane >= 0; plane--) /* byte-plane transposed copy, :143-155 */
        for (int i = plane; i < L; i += 4) t[idx++] = x[i];
      for (int i = 0; i < L; i++) out[i] = (i == p->root) ? t[i] : t[i - 1]; /* :157-171; root must be 0 */
      break;
    }
    case ORC_PRED_DIFF:
      for (int i = 0; i < L; i++)
        out[i] = (i == p->root) ? x[i] : (uint8_t)((uint8_t)p->diff[i] + x[p->base[i]]);
      break;
    case ORC_PRED_WEIGHT:
      for (int i = 0; i < L; i++) {
        if (i == p->root) { out[i] = x[i]; continue; }
        int b = x[p->base[i]], s = p->shift[i];
        int a = s < 0 ? -s : s;
        int v = (a >= 24) ? 0 : (s < 0 ? (b >> a) : (b << a));
        out[i] = (uint8_t)v;
      }
      break;
  }
}

/* PredCompModule::CompressLine, PredCompModule.cpp:11-26: residue -> bitplane -> xor -> scan.
 * S receives 8L bits (one byte per bit) in scan order; returns the number of leading all-zero rows. */
static int predcomp_scan(const orc_predcomp* p, const uint8_t* x, int L, uint8_t* S, uint8_t* pred_out) {
  uint8_t pred[ORC_MAX_LINE], r[ORC_MAX_LINE], g[ORC_MAX_LINE];
  predict(p, x, L, pred);
  if (pred_out) memcpy(pred_out, pred, (size_t)L);
  /* ResidueModule::ProcessLine, ResidueModule.cpp:12-41: root first, then the others in order */
  r[0] = x[p->root];
  int j = 1;
  for (int i = 0; i < L; i++)
    if (i != p->root) r[j++] = (uint8_t)(x[i] - pred[i]);
  /* BitplaneModule (plane b = bit 7-b, BitplaneModule.cpp:25-36) + XORModule (XORModule.cpp:9-20),
   * folded per byte: planes b>=1 of columns k>=1 are XORed with plane b-1 (consecutive) or plane 0. */
  g[0] = r[0];
  for (int k = 1; k < L; k++)
    g[k] = p->consecutive_xor ? (uint8_t)(r[k] ^ (r[k] >> 1)) : (uint8_t)(r[k] ^ ((r[k] & 0x80) ? 0x7f : 0));
  /* ScanModule::ProcessLine, ScanModule.cpp:6-22 */
  int nbits = 8 * L;
  memset(S, 0, (size_t)nbits);
  for (int i = 0; i < p->table_size; i++) S[i] = (g[p->cols[i]] >> (7 - p->rows[i])) & 1;
  /* leading zero rows, VPC.cpp:378-387 */
  int z = 0;
  for (int row = 0; row < nbits / 16; row++) {
    int nz = 0;
    for (int k = 0; k < 16; k++) nz |= S[row * 16 + k];
    if (nz) break;
    z++;
  }
  return z;
}

/* FPCModule::ProcessLine + row tests, FPCModule.cpp:19-85, 87-158; costs FPCModule.h:55 */
static int common_encoder(const uint8_t* S, int nrows) {
  int cost = 0, run = 0;
  for (int row = 0; row < nrows; row++) {
    const uint8_t* v = S + row * 16;
    int ones = 0, first = -1, second = -1, front = 0, back = 0;
    for (int k = 0; k < 16; k++)
      if (v[k]) {
        if (ones == 0) first = k; else if (ones == 1) second = k;
        ones++;
        if (k < 8) front = 1; else back = 1;
      }
    if (ones == 0) { run++; continue; }
    if (run > 0) { cost += (run > 1) ? 7 : 4; run = 0; }
    if (ones == 1) cost += 7;
    else if (ones == 2 && second - first == 1) cost += 8;
    else if (!front) cost += 12;
    else if (!back) cost += 12;
    else cost += 17;
  }
  if (run > 0) cost += (run > 1) ? 7 : 4;
  return cost;
}

/* One block.  Returns size in bits incl. encoding bits; *sel = cluster (-1 uncompressed).
 * abs_sum/sq_sum (may be null) receive sum|r'| and sum r'^2 over all L bytes for blocks that reach
 * the predictor stage (VPC.cpp:417-443, ResidueModule.cpp:43-73); *stage3 = 1 for those blocks. */
unsigned orc_compress_block(const orc_config* c, const uint8_t* x, int* sel, uint64_t* abs_sum,
                            uint64_t* sq_sum, int* stage3) {
  const int L = c->line_size;
  if (stage3) *stage3 = 0;
  /* AllZeroModule.cpp:7-15 + VPC::checkAllZeros VPC.cpp:332-347 */
  int allzero = 1;
  for (int i = 0; i < L; i++) if (x[i]) { allzero = 0; break; }
  if (allzero) { *sel = 0; return (unsigned)c->enc_bits[1]; }
  /* AllWordSameModule.cpp:7-21 + VPC::checkAllWordSame VPC.cpp:349-364 */
  if (c->has_wordsame) {
    int same = 1;
    for (int i = 4; i < L; i++) if (x[i] != x[i % 4]) { same = 0; break; }
    if (same) { *sel = 1; return 32u + (unsigned)c->enc_bits[2]; }
  }
  /* VPC::checkOtherPatterns VPC.cpp:366-415 */
  uint8_t S[8 * ORC_MAX_LINE], bestS[8 * ORC_MAX_LINE], pred[ORC_MAX_LINE], bestpred[ORC_MAX_LINE];
  int best = -1, bestz = 0;
  memset(bestS, 0, sizeof(bestS));
  for (int m = 0; m < c->num_predcomp; m++) {
    int z = predcomp_scan(&c->pc[m], x, L, S, pred);
    if (bestz <= z) { /* ties -> later module, VPC.cpp:389 */
      best = c->first_predcomp + m;
      bestz = z;
      memcpy(bestS, S, (size_t)8 * L);
      memcpy(bestpred, pred, (size_t)L);
    }
  }
  int cost = (best >= 0) ? common_encoder(bestS, 8 * L / 16) : 0; /* empty Binary -> 0, FPCModule.cpp:25 */
  unsigned size;
  if (cost < 8 * L) size = (unsigned)cost; else { best = -1; size = 8u * (unsigned)L; }
  size += (unsigned)c->enc_bits[best + 1];
  *sel = best;
  uint64_t a = 0, s = 0;
  for (int i = 0; i < L; i++) {
    uint8_t rr = (best >= 0) ? (uint8_t)(x[i] - bestpred[i]) : x[i];
    a += rr;
    s += (uint64_t)rr * rr;
  }
  if (abs_sum) *abs_sum = a;
  if (sq_sum) *sq_sum = s;
  if (stage3) *stage3 = 1;
  return size;
}

/* Stats layout (uint64), clusters indexed by sel+1 in [0, N]:
 *   stats[0]=blocks stats[1]=OriginalSize stats[2]=CompressedSize   (CompResult.h:30-35)
 *   stats[3 + 5*k + {0..4}] = count, compressedSize, stage-3 lines, sum|r|, sum r^2   (VPC.h:49-76)
 *   hist[k*hist_bins + size]                                                       (VPC.h:58)  */
typedef struct {
  const orc_config* c;
  const uint8_t* lines;
  uint64_t lo, hi;
  uint32_t* sizes;
  int32_t* sels;
  uint64_t* stats;
  uint64_t* hist;
  unsigned hist_bins;
} orc_job;

static void* orc_worker(void* arg) {
  orc_job* j = (orc_job*)arg;
  const orc_config* c = j->c;
  const int L = c->line_size;
  for (uint64_t i = j->lo; i < j->hi; i++) {
    int sel, st3;
    uint64_t a = 0, s = 0;
    unsigned size = orc_compress_block(c, j->lines + (size_t)i * L, &sel, &a, &s, &st3);
    if (j->sizes) j->sizes[i] = size;
    if (j->sels) j->sels[i] = sel;
    int k = sel + 1;
    uint64_t* st = j->stats;
    st[0]++; st[1] += 8u * L; st[2] += size;
    st[3 + 5 * k + 0]++;
    st[3 + 5 * k + 1] += size;
    if (st3) { st[3 + 5 * k + 2]++; st[3 + 5 * k + 3] += a; st[3 + 5 * k + 4] += s; }
    if (j->hist && size < j->hist_bins) j->hist[(size_t)k * j->hist_bins + size]++;
  }
  return NULL;
}

/* threads <= 1 runs inline (the reference itself is single-threaded, main.cpp:229-244) */
void orc_run(const orc_config* c, const uint8_t* lines, uint64_t n, uint32_t* sizes, int32_t* sels,
             uint64_t* stats, uint64_t* hist, unsigned hist_bins, int threads) {
  const int K = c->num_modules + 1;
  const int nst = 3 + 5 * K;
  if (threads < 1) threads = 1;
  if ((uint64_t)threads > n) threads = n ? (int)n : 1;
  orc_job* jobs = (orc_job*)calloc((size_t)threads, sizeof(orc_job));
  pthread_t* tid = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  for (int t = 0; t < threads; t++) {
    orc_job* j = &jobs[t];
    j->c = c; j->lines = lines; j->sizes = sizes; j->sels = sels; j->hist_bins = hist_bins;
    j->lo = n * (uint64_t)t / (uint64_t)threads;
    j->hi = n * (uint64_t)(t + 1) / (uint64_t)threads;
    j->stats = (uint64_t*)calloc((size_t)nst, sizeof(uint64_t));
    j->hist = hist ? (uint64_t*)calloc((size_t)K * hist_bins, sizeof(uint64_t)) : NULL;
    if (threads > 1) pthread_create(&tid[t], NULL, orc_worker, j); else orc_worker(j);
  }
  for (int t = 0; t < threads; t++) {
    orc_job* j = &jobs[t];
    if (threads > 1) pthread_join(tid[t], NULL);
    if (stats) for (int q = 0; q < nst; q++) stats[q] += j->stats[q];
    if (hist) for (size_t q = 0; q < (size_t)K * hist_bins; q++) hist[q] += j->hist[q];
    free(j->stats);
    free(j->hist);
  }
  free(jobs);
  free(tid);
}
