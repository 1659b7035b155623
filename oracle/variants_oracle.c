/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the secondary compressor variants (BDI, FPC, BPC).
 * Filled in together with the corresponding CUDA kernels; see mpc_oracle.c for the header rules.
 */
#include <stdint.h>
int orc_variants_placeholder(void) { return 0; }
