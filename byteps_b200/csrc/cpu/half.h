// Scalar fp16 / bf16 <-> fp32 conversions (round-to-nearest-even) for host code.
// The reference vendors mshadow's half.h (/root/reference/byteps/common/half.h);
// this is an independent minimal implementation that also covers bf16.
#pragma once
#include <cstdint>
#include <cstring>

namespace bps {

inline float bf16_to_f32(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x0040u);  // quiet NaN
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return static_cast<uint16_t>(u >> 16);
}

inline float f16_to_f32(uint16_t h) {
  uint32_t sign = (h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t u;
  if (exp == 0) {
    if (man == 0) {
      u = sign;
    } else {  // subnormal
      int e = -1;
      do {
        man <<= 1;
        ++e;
      } while (!(man & 0x400u));
      man &= 0x3ffu;
      u = sign | ((127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    u = sign | 0x7f800000u | (man << 13);
  } else {
    u = sign | ((exp + 112) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &u, 4);
  return f;
}

inline uint16_t f32_to_f16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t sign = (u >> 16) & 0x8000u;
  uint32_t au = u & 0x7fffffffu;
  if (au >= 0x7f800000u) return static_cast<uint16_t>(sign | 0x7c00u | (au > 0x7f800000u ? 0x200u : 0));
  if (au >= 0x477ff000u) return static_cast<uint16_t>(sign | 0x7c00u);  // overflow -> inf
  if (au < 0x33000001u) return static_cast<uint16_t>(sign);              // underflow -> 0
  int32_t exp = static_cast<int32_t>(au >> 23) - 127 + 15;
  uint32_t man = au & 0x7fffffu;
  if (exp <= 0) {  // subnormal half
    man |= 0x800000u;
    int shift = 14 - exp;
    uint32_t half_man = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1);
    uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half_man & 1u))) ++half_man;
    return static_cast<uint16_t>(sign | half_man);
  }
  uint32_t half = (static_cast<uint32_t>(exp) << 10) | (man >> 13);
  uint32_t rem = man & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) ++half;
  return static_cast<uint16_t>(sign | half);
}

}  // namespace bps
