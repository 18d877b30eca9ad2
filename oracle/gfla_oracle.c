/*
 * TEST INFRASTRUCTURE ONLY -- see gfla_oracle_impl.h for the contract.
 * Builds liboracle.so: the float and double instantiations of the CPU
 * restatement of the GFLA warping path.  Parity status: PINNED -- checked
 * bit-for-bit (single host thread) against the reference's own kernel bodies
 * compiled for the host (oracle/_ref, tests/test_oracle_vs_ref.py runs where
 * /root/reference exists) and against the committed vectors in tests/golden/
 * that were generated from those reference bodies.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SFX(x) x##_f32
#define FLOOR(v) floorf(v)
#define EXP(v) expf(v)
#include "gfla_oracle_impl.h"
#undef REAL
#undef SFX
#undef FLOOR
#undef EXP

#define REAL double
#define SFX(x) x##_f64
#define FLOOR(v) floor(v)
#define EXP(v) exp(v)
#include "gfla_oracle_impl.h"
#undef REAL
#undef SFX
#undef FLOOR
#undef EXP

int oracle_abi_version(void) { return 1; }
