// Host instances of the DEFLATE decoder (inflate_fast.hpp): one compiled for the baseline x86-64 ISA and one for BMI2 (the
// decoder is all variable shifts and bit-field extracts, which are single-uop SHRX / SHLX / BZHI there), picked once at run time.
#include "inflate_fast.hpp"

namespace strling {

namespace {

int inflate_generic(infl::Tables &T, const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len) {
  return infl::inflate_block(T, in, in_len, out, out_len);
}

#if defined(__x86_64__)
__attribute__((target("bmi2,bmi"))) int inflate_bmi2(infl::Tables &T, const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len) {
  return infl::inflate_block(T, in, in_len, out, out_len);
}
#endif

}  // namespace

int inflate_block_host(infl::Tables &T, const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len) {
#if defined(__x86_64__)
  static const bool bmi2 = __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("bmi");
  if (bmi2) return inflate_bmi2(T, in, in_len, out, out_len);
#endif
  return inflate_generic(T, in, in_len, out, out_len);
}

}  // namespace strling
