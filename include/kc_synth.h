/* kc_synth.h — the frozen synthetic protein-set generator "G1" used by the benchmarks and tests
 * (BASELINE.md records the seeds).  Host only, plain C ABI, its own small library
 * (libkc_synth.so, csrc/synth.cpp): not part of the engine's drop-in boundary. */
#ifndef KC_SYNTH_H_
#define KC_SYNTH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Generator "G1" (frozen; BASELINE.md / DESIGN.md give the law).  All-integer and
 * counter-based, so any subset of proteins can be generated independently and in parallel.
 *   length_law 0 ("A"): 50 + sum of four uniform ints in [0,150]   (mean 350)
 *   length_law 1 ("B"): 50 + 1950 * (t / 2^16)^4, t uniform in [0, 65535] (50..2000, skewed)
 * Families of 16 consecutive proteins share a base sequence; member j re-draws each residue
 * with probability j * 1311 / 65536.  class = family % 15, except every 8th family where
 * class = (family + j) % 15.
 * Step 1 fills offsets[n+1] and class_id[n]; step 2 fills residues[offsets[n]]. */
int kc_synth_layout(uint64_t n, int length_law, uint64_t seed, uint64_t* offsets, uint32_t* class_id);
int kc_synth_residues(uint64_t n, int length_law, uint64_t seed, int threads, const uint64_t* offsets,
                      uint8_t* residues);

#ifdef __cplusplus
}
#endif
#endif /* KC_SYNTH_H_ */
