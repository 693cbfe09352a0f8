// synth.cpp — the frozen synthetic protein-set generator "G1" (include/kc_synth.h): benchmark and test
// tooling, host only (g++, no CUDA), built as libkc_synth.so so that nothing that only needs a protein
// set (bench.py --impl reference, the golden generators) has to load the engine library.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/kc_synth.h"

namespace {

template <class F>
void parallel_chunks(int threads, uint64_t n, uint64_t grain, F f) {
  if (threads <= 1 || n <= grain) {
    f(0, n);
    return;
  }
  std::atomic<uint64_t> cursor{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] {
      for (;;) {
        const uint64_t lo = cursor.fetch_add(grain);
        if (lo >= n) break;
        f(lo, std::min(n, lo + grain));
      }
    });
  for (auto& th : pool) th.join();
}

// ---- generator G1 ------------------------------------------------------------------------
struct SplitMix {
  uint64_t s;
  uint64_t next() {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
};

SplitMix stream(uint64_t seed, uint64_t index, uint64_t tag) {
  SplitMix g{seed ^ ((index + 1) * 0x9E3779B97F4A7C15ull) ^ (tag * 0xC2B2AE3D27D4EB4Full)};
  g.next();
  return g;
}

// residue frequencies of the ARG protein set, total 3 436 746
const char kLetters[21] = "LAGVISTFREKDPQNYMHWC";
const uint32_t kCounts[20] = {377380, 336508, 269059, 258845, 257326, 206108, 194902, 176903, 163437, 158642,
                              158511, 153089, 143715, 124590, 118853, 101563, 92047,  64642,  54450,  26176};
const uint32_t kTotal = 3436746;

struct Cumulative {
  uint32_t c[20];
  Cumulative() {
    uint32_t s = 0;
    for (int i = 0; i < 20; ++i) {
      s += kCounts[i];
      c[i] = s;
    }
  }
  char pick(uint32_t u) const {  // u uniform in [0, kTotal)
    int lo = 0, hi = 19;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (u < c[mid]) hi = mid; else lo = mid + 1;
    }
    return kLetters[lo];
  }
};
const Cumulative kCum;

uint32_t family_length(uint64_t seed, uint64_t fam, int law) {
  SplitMix g = stream(seed, fam, 1);
  if (law == 0) {
    uint32_t len = 50;
    for (int i = 0; i < 4; ++i) len += (uint32_t)(((g.next() >> 32) * 151ull) >> 32);
    return len;
  }
  const uint64_t t = ((g.next() >> 32) * 65536ull) >> 32;
  const uint64_t t4 = t * t * t * t;
  return 50u + (uint32_t)(((t4 >> 32) * 1950ull) >> 32);
}

}  // namespace

extern "C" {

int kc_synth_layout(uint64_t n, int law, uint64_t seed, uint64_t* offsets, uint32_t* class_id) {
  if (!offsets || (n && !class_id) || (law != 0 && law != 1)) return 1;
  offsets[0] = 0;
  uint32_t len = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t fam = i / 16, j = i % 16;
    if (j == 0 || i == 0) len = family_length(seed, fam, law);
    offsets[i + 1] = offsets[i] + len;
    class_id[i] = (uint32_t)((fam % 8 == 7) ? (fam + j) % 15 : fam % 15);
  }
  return 0;
}

int kc_synth_residues(uint64_t n, int law, uint64_t seed, int threads, const uint64_t* offsets,
                      uint8_t* residues) {
  if (!offsets || (n && !residues) || (law != 0 && law != 1)) return 1;
  (void)law;
  const uint64_t n_fam = (n + 15) / 16;
  parallel_chunks(threads < 1 ? 1 : threads, n_fam, 64, [&](uint64_t lo, uint64_t hi) {
    std::vector<uint8_t> base;
    for (uint64_t fam = lo; fam < hi; ++fam) {
      const uint64_t first = fam * 16;
      const uint64_t len = offsets[first + 1] - offsets[first];
      base.resize(len);
      SplitMix g = stream(seed, fam, 2);
      for (uint64_t p = 0; p < len; ++p) {
        const uint64_t r = g.next();
        base[p] = (r & 8191u) == 0 ? (uint8_t)'X' : (uint8_t)kCum.pick((uint32_t)(((r >> 32) * kTotal) >> 32));
      }
      for (uint64_t j = 0; j < 16 && first + j < n; ++j) {
        uint8_t* dst = residues + offsets[first + j];
        SplitMix m = stream(seed, first + j, 3);
        const uint32_t thr = (uint32_t)j * 1311u;
        for (uint64_t p = 0; p < len; ++p) {
          const uint64_t r = m.next();
          dst[p] = (uint32_t)(r & 0xFFFFu) < thr
                       ? (uint8_t)kCum.pick((uint32_t)((((r >> 16) & 0xFFFFFFFFull) * kTotal) >> 32))
                       : base[p];
        }
      }
    }
  });
  return 0;
}

}  // extern "C"
