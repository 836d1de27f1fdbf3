// orb_ref.cpp — CPU oracle pieces for the ORBextractor stages.
// TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may
// load this; the product path (lorb_slam_b200/) never does.
//
// Part 1: pins of the device's sinf/cosf restatement.  computeOrbDescriptor (reference
// src/ORBextractor.cpp:116) calls the C library's cosf/sinf; lorb_slam_b200/csrc/libm_sincosf.cuh
// restates glibc's algorithm for the device.  Its host instantiation is compared here with the
// libm this process links, bit for bit.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../lorb_slam_b200/csrc/libm_sincosf.cuh"

extern "C" {

void orc_libm_sincosf(int n, const float* x, float* s, float* c) {
  for (int i = 0; i < n; i++) {
    s[i] = sinf(x[i]);
    c[i] = cosf(x[i]);
  }
}

void orc_restated_sincosf(int n, const float* x, float* s, float* c) {
  for (int i = 0; i < n; i++) {
    s[i] = lorb_libm::sinf_libm(x[i]);
    c[i] = lorb_libm::cosf_libm(x[i]);
  }
}

// Number of floats with bit patterns lo, lo+stride, ... <= hi (and their negatives) on which the
// restatement and libm disagree in sinf or cosf.
long long orc_sincosf_mismatches(uint32_t lo, uint32_t hi, uint32_t stride) {
  long long bad = 0;
  const long long cnt = ((long long)hi - lo) / stride + 1;
#pragma omp parallel for reduction(+ : bad) schedule(static)
  for (long long k = 0; k < cnt; k++) {
    const uint32_t bits = lo + (uint32_t)k * stride;
    for (int neg = 0; neg < 2; neg++) {
      const uint32_t b = bits | (neg ? 0x80000000u : 0u);
      float x;
      memcpy(&x, &b, 4);
      const float s = sinf(x), c = cosf(x), s2 = lorb_libm::sinf_libm(x), c2 = lorb_libm::cosf_libm(x);
      if (memcmp(&s, &s2, 4) || memcmp(&c, &c2, 4)) bad++;
    }
  }
  return bad;
}

}  // extern "C"
