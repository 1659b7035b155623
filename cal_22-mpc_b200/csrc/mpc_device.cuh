// Device-side building blocks shared by the MPC kernels (sm_100a).
//
// Everything here is integer SWAR work on packed bytes / packed 16-bit scan rows.  The
// functions are __host__ __device__ so that tests/ can compile this header with g++ and check
// the bit tricks exhaustively on the CPU (tests/test_swar_host.py); on the device the
// intrinsics map to PRMT / LOP3 / VIMNMX.U16x2 / VABSDIFF4 / IDP.4A.
#pragma once
#if defined(__CUDACC_RTC__)
// NVRTC has no host headers: fixed-width types by hand (LP64)
typedef unsigned char uint8_t;
typedef signed char int8_t;
typedef unsigned short uint16_t;
typedef short int16_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
typedef unsigned long size_t;
#else
#include <stdint.h>
#endif

#if defined(__CUDACC__)
#define MPC_HD __host__ __device__ __forceinline__
#define MPC_HDM __host__ __device__ __forceinline__  // member functions
#else
#define MPC_HD static inline
#define MPC_HDM inline
#endif

namespace mpcdev {

// ---- packed-halfword min (DPX on sm_90+/sm_100: VIMNMX.U16x2 / VIMNMX3.U16x2) -------------------
MPC_HD uint32_t min_u16x2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __vminu2(a, b);
#else
  uint32_t lo = (a & 0xffffu) < (b & 0xffffu) ? (a & 0xffffu) : (b & 0xffffu);
  uint32_t hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
  return lo | (hi << 16);
#endif
}
MPC_HD uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  return __vimin3_u16x2(a, b, c);
#else
  return min_u16x2(min_u16x2(a, b), c);
#endif
}

// packed-halfword add, wrapping per halfword (PTX add.u16x2 -> VIADD.U16x2 on sm_90+/sm_100)
MPC_HD uint32_t add_u16x2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("add.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
#else
  return ((a + b) & 0xffffu) | ((((a >> 16) + (b >> 16)) & 0xffffu) << 16);
#endif
}

// x >> n issued on the FMA pipe (IMAD.HI) instead of the ALU pipe (SHF): the kernels are ALU-pipe bound
MPC_HD uint32_t shr_fma(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(1u << (32 - n)));
  return d;
#else
  return x >> n;
#endif
}
// a * k + b on the FMA pipe; `volatile` keeps ptxas from folding the result back into ALU-pipe forms (LEA.HI, IADD3)
MPC_HD uint32_t mad_fma(uint32_t a, uint32_t k, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(k), "r"(b));
  return d;
#else
  return a * k + b;
#endif
}
MPC_HD uint32_t shr_fma_opaque(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(1u << (32 - n)));
  return d;
#else
  return x >> n;
#endif
}
// a - b issued on the FMA pipe (IMAD) instead of the ALU pipe (IADD3)
MPC_HD uint32_t sub_fma(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(d) : "r"(b), "r"(a));
  return d;
#else
  return a - b;
#endif
}

// ---- per-byte modular arithmetic on 4 packed bytes -----------------------------------------------
// (a - b) mod 256 per byte.  ResidueModule::ProcessLine, ResidueModule.cpp:34.
MPC_HD uint32_t sub_u8x4(uint32_t a, uint32_t b) {
  const uint32_t H = 0x80808080u;
  uint32_t t = sub_fma(a | H, b & ~H);
  return t ^ ((a ^ ~b) & H);
}
// (a + b) mod 256 per byte.  DiffBasePredictor::PredictLine, PredictorModule.cpp:104.
MPC_HD uint32_t add_u8x4(uint32_t a, uint32_t b) {
  const uint32_t H = 0x80808080u;
  uint32_t t = (a & ~H) + (b & ~H);
  return t ^ ((a ^ b) & H);
}

// ---- (a - b) mod 256 per byte in 16-bit lanes, for predicted bytes that are SHIFTED line bytes (WeightBasePredictor) ----------
// sub_u8x4 costs four ALU-pipe instructions per word plus whatever assembles b (masks of the shifted bytes).  Here the line word
// is split once into its even bytes [a0, 1, a2, 0] and its odd bytes [0, a1, 1, a3] (one PRMT each; the 1s are borrow guards),
// every predicted byte is SUBTRACTED in place by a multiply-add on the FMA pipe (value * -2^k + lanes: the multiplier is the shift
// and the byte position at once), and one LOP3 puts the two halves together.  A term must be exact in its own byte and zero below
// it; what it leaves above the byte is harmless in byte 3 of the even half (discarded) and beyond bit 31.
MPC_HD uint32_t lanes_even(uint32_t a) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, 1u, 0x5240u);
#else
  return (a & 0x00ff00ffu) | 0x00000100u;
#endif
}
MPC_HD uint32_t lanes_odd(uint32_t a) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, 1u, 0x3415u);
#else
  return (a & 0xff00ff00u) | 0x00010000u;
#endif
}
MPC_HD uint32_t lanes_merge(uint32_t even, uint32_t odd) {
#if defined(__CUDA_ARCH__)
  uint32_t d;  // one bit select (written with two masks ptxas emits two LOP3)
  asm("lop3.b32 %0, %1, %2, 0x00ff00ff, 0xe4;" : "=r"(d) : "r"(even), "r"(odd));
  return d;
#else
  return (even & 0x00ff00ffu) | (odd & 0xff00ff00u);
#endif
}

// ---- bit-plane XOR folded into the byte domain ------------------------------------------------------
// BitplaneModule (plane b = bit 7-b of the residue byte, BitplaneModule.cpp:25-36) followed by
// XORModule (XORModule.cpp:9-20).  `keep` has 0xff in every byte lane that must stay untouched
// (column 0 of the residue line, XORModule.cpp:12 "j = 1").
MPC_HD uint32_t xor_planes_consecutive(uint32_t r, uint32_t keep) {
  return r ^ ((r >> 1) & 0x7f7f7f7fu & ~keep);
}
MPC_HD uint32_t xor_planes_first(uint32_t r, uint32_t keep) {
  uint32_t msb = (r >> 7) & 0x01010101u & ~keep;
  return r ^ (msb * 0x7fu);
}

// ---- common encoder on two packed scan rows -----------------------------------------------------------
// A scan row is 16 bits; bit 15 holds scan position 0 (so the "front half" is the high byte).
// w carries two rows (low / high halfword).  Returns, per halfword, the cost in bits of a
// NON-ZERO row (FPCModule.cpp:47-66, costs FPCModule.h:55: single one 7, two consecutive ones 8,
// front or back half zero 12, else 17) and 0 for an all-zero row; *nz gets 1 per non-zero row.
MPC_HD uint32_t row2_cost(uint32_t w, uint32_t* nz_out) {
  const uint32_t ONE = 0x00010001u;
  uint32_t nz = min_u16x2(w, ONE);
  uint32_t t = add_u16x2(w, 0xffffffffu);  // row - 1 per halfword (no borrow across the halves)
  uint32_t s = w & t;                 // row with its lowest set bit cleared: != 0 <=> two or more ones
  uint32_t lb = w & ~t;               // lowest set bit (0 for a zero row)
  uint32_t x = (shr_fma(s, 1) & 0x7fff7fffu) ^ lb;  // == 0 <=> exactly two ones, adjacent (or fewer than two ones)
  uint32_t back = w & 0x00ff00ffu;
  uint32_t front = shr_fma(w, 8) & 0x00ff00ffu;
  uint32_t a = min3_u16x2(s, x, ONE);          // 1 <=> >= 2 ones and not "two consecutive"
  uint32_t b3 = min3_u16x2(a, front, back);    // 1 <=> additionally both halves non-zero
  uint32_t ns = min_u16x2(s, ONE);
  *nz_out = nz;
  return nz * 7u + ns + a * 4u + b3 * 5u;
}

// Cost of the zero rows of a block given the 64-bit mask of all-zero rows in scan order (bit i = row i,
// only the low `rows` bits meaningful).  A maximal run of >= 2 zero rows costs 7, an isolated one 4
// (FPCModule.cpp:27-45, 69-79).
MPC_HD uint32_t zero_run_cost(uint64_t zmask) {
  uint64_t start = zmask & ~(zmask << 1);
  uint64_t single = start & ~(zmask >> 1);
#if defined(__CUDA_ARCH__)
  uint32_t ns = (uint32_t)__popcll(start), n1 = (uint32_t)__popcll(single);
#else
  uint32_t ns = (uint32_t)__builtin_popcountll(start), n1 = (uint32_t)__builtin_popcountll(single);
#endif
  return 7u * ns - 3u * n1;
}

// Number of leading all-zero rows (VPC.cpp:378-387); `rows` <= 64.
MPC_HD uint32_t leading_zero_rows(uint64_t zmask, uint32_t rows) {
  uint64_t nzm = ~zmask;
  if (rows < 64) nzm |= ~0ull << rows;
#if defined(__CUDA_ARCH__)
  return nzm ? (uint32_t)(__ffsll((long long)nzm) - 1) : rows;
#else
  return nzm ? (uint32_t)__builtin_ctzll(nzm) : rows;
#endif
}

// sum of the 4 bytes / sum of the squares of the 4 bytes (MAE / MSE numerators, ResidueModule.cpp:43-73)
MPC_HD uint32_t sum_u8x4(uint32_t r) {
#if defined(__CUDA_ARCH__)
  return __vsadu4(r, 0u);
#else
  return (r & 0xff) + ((r >> 8) & 0xff) + ((r >> 16) & 0xff) + (r >> 24);
#endif
}
// acc + sum of the 4 bytes as a dot product with ones (IDP.4A), off the ALU pipe
MPC_HD uint32_t sum_u8x4_acc(uint32_t r, uint32_t acc) {
#if defined(__CUDA_ARCH__)
  return __dp4a(r, 0x01010101u, acc);
#else
  return acc + (r & 0xff) + ((r >> 8) & 0xff) + ((r >> 16) & 0xff) + (r >> 24);
#endif
}
MPC_HD uint32_t sumsq_u8x4(uint32_t r) {
#if defined(__CUDA_ARCH__)
  return __dp4a(r, r, 0u);
#else
  uint32_t s = 0;
  for (int k = 0; k < 4; k++) { uint32_t b = (r >> (8 * k)) & 0xff; s += b * b; }
  return s;
#endif
}

// splitmix64 finaliser used by the synthetic dump generator (tools/gen_dump.py has the same function)
MPC_HD uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace mpcdev
