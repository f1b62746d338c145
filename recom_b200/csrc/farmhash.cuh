// farmhash.cuh — device Fingerprint64 (farmhashna::Hash64), written for registers.
//
// Semantics: TF StringToHashBucketFast hashes with Fingerprint64 ==
// farmhashna::Hash64 (reference tree: examples/cc/tensorflow-v2.6.2/
// tensorflow/core/platform/fingerprint.h:80-91; algorithm pinned by
// third_party/farmhash/workspace.bzl:9-10, google/farmhash @816a4ae).
//
// The hash is templated on a byte accessor so the same code serves
//   * MemBytes : a string that lives in the request blob (HASH_STR), and
//   * RegBytes : a <=24-byte string held in three 64-bit registers, used for the
//                decimal rendering of an integer (HASH_INT = AsString + hash,
//                feature_column_v2.py:3414-3419) without touching local memory.
#pragma once
#include <stdint.h>

namespace recom {
namespace fh {

constexpr uint64_t kK0 = 0xc3a5c85c97cb3127ULL;
constexpr uint64_t kK1 = 0xb492b66fbe98f273ULL;
constexpr uint64_t kK2 = 0x9ae16a3b2f90404fULL;

__device__ __forceinline__ uint64_t rot(uint64_t v, int s) {
  return (v >> s) | (v << (64 - s));  // s is never 0 at the call sites
}
__device__ __forceinline__ uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }

__device__ __forceinline__ uint64_t hash_len16(uint64_t u, uint64_t v, uint64_t mul) {
  uint64_t a = (u ^ v) * mul;
  a ^= (a >> 47);
  uint64_t b = (v ^ a) * mul;
  b ^= (b >> 47);
  return b * mul;
}

// Bytes in global memory at an arbitrary alignment. fetch32/fetch64 read the aligned
// words that cover the requested bytes and funnel-shift them: every word read holds at
// least one byte of the string, so reads stay inside the string's slot as long as the
// slot is padded to an 8-byte boundary (include/recom_b200.h: blob slots are padded).
struct MemBytes {
  const uint8_t* p;
  __device__ __forceinline__ uint32_t byte(uint32_t pos) const { return __ldg(p + pos); }
  __device__ __forceinline__ uint32_t fetch32(uint32_t pos) const {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) + pos;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t lo = __ldg(w);
    if (sh == 0) return lo;
    return __funnelshift_r(lo, __ldg(w + 1), sh);
  }
  __device__ __forceinline__ uint64_t fetch64(uint32_t pos) const {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) + pos;
    const unsigned long long* w =
        reinterpret_cast<const unsigned long long*>(a & ~uintptr_t(7));
    const uint32_t sh = (uint32_t)(a & 7) * 8;
    const uint64_t lo = __ldg(w);
    if (sh == 0) return lo;
    return (lo >> sh) | ((uint64_t)__ldg(w + 1) << (64 - sh));
  }
};

// Up to 24 bytes packed little-endian into w0 (bytes 0-7), w1 (8-15), w2 (16-23).
struct RegBytes {
  uint64_t w0, w1, w2;
  __device__ __forceinline__ uint64_t fetch64(uint32_t pos) const {
    const uint32_t word = pos >> 3, sh = (pos & 7) * 8;
    const uint64_t lo = word == 0 ? w0 : (word == 1 ? w1 : w2);
    const uint64_t hi = word == 0 ? w1 : (word == 1 ? w2 : 0ull);
    return sh == 0 ? lo : ((lo >> sh) | (hi << (64 - sh)));
  }
  __device__ __forceinline__ uint32_t fetch32(uint32_t pos) const {
    return (uint32_t)fetch64(pos);
  }
  __device__ __forceinline__ uint32_t byte(uint32_t pos) const {
    return (uint32_t)fetch64(pos) & 0xffu;
  }
};

// HashLen0to16 for 8 <= len <= 16; first = bytes [0, 8), last = bytes [len - 8, len).
__device__ __forceinline__ uint64_t hash_len_8_16(uint64_t first, uint64_t last, uint32_t len) {
  const uint64_t mul = kK2 + (uint64_t)len * 2;
  const uint64_t a = first + kK2;
  const uint64_t b = last;
  const uint64_t c = rot(b, 37) * mul + a;
  const uint64_t d = (rot(a, 25) + b) * mul;
  return hash_len16(c, d, mul);
}

template <class B>
__device__ __forceinline__ uint64_t hash_len_0_16(const B& s, uint32_t len) {
  if (len >= 8) return hash_len_8_16(s.fetch64(0), s.fetch64(len - 8), len);
  if (len >= 4) {
    const uint64_t mul = kK2 + (uint64_t)len * 2;
    const uint64_t a = s.fetch32(0);
    return hash_len16((uint64_t)len + (a << 3), s.fetch32(len - 4), mul);
  }
  if (len > 0) {
    const uint32_t a = s.byte(0), b = s.byte(len >> 1), c = s.byte(len - 1);
    const uint32_t y = a + (b << 8);
    const uint32_t z = len + (c << 2);
    return shift_mix((uint64_t)y * kK2 ^ (uint64_t)z * kK0) * kK2;
  }
  return kK2;
}

template <class B>
__device__ __forceinline__ uint64_t hash_len_17_32(const B& s, uint32_t len) {
  const uint64_t mul = kK2 + (uint64_t)len * 2;
  const uint64_t a = s.fetch64(0) * kK1;
  const uint64_t b = s.fetch64(8);
  const uint64_t c = s.fetch64(len - 8) * mul;
  const uint64_t d = s.fetch64(len - 16) * kK2;
  return hash_len16(rot(a + b, 43) + rot(c, 30) + d, a + rot(b + kK2, 18) + c, mul);
}

struct Pair {
  uint64_t first, second;
};

__device__ __forceinline__ Pair weak32(uint64_t w, uint64_t x, uint64_t y, uint64_t z,
                                       uint64_t a, uint64_t b) {
  a += w;
  b = rot(b + a + z, 21);
  const uint64_t c = a;
  a += x;
  a += y;
  b += rot(a, 44);
  return Pair{a + z, b + c};
}

__device__ __forceinline__ Pair weak32_at(const MemBytes& s, uint32_t pos, uint64_t a,
                                          uint64_t b) {
  return weak32(s.fetch64(pos), s.fetch64(pos + 8), s.fetch64(pos + 16),
                s.fetch64(pos + 24), a, b);
}

__device__ __forceinline__ uint64_t hash_len_33_64(const MemBytes& s, uint32_t len) {
  const uint64_t mul = kK2 + (uint64_t)len * 2;
  const uint64_t a = s.fetch64(0) * kK2;
  const uint64_t b = s.fetch64(8);
  const uint64_t c = s.fetch64(len - 8) * mul;
  const uint64_t d = s.fetch64(len - 16) * kK2;
  const uint64_t y = rot(a + b, 43) + rot(c, 30) + d;
  const uint64_t z = hash_len16(y, a + rot(b + kK2, 18) + c, mul);
  const uint64_t e = s.fetch64(16) * mul;
  const uint64_t f = s.fetch64(24);
  const uint64_t g = (y + s.fetch64(len - 32)) * mul;
  const uint64_t h = (z + s.fetch64(len - 24)) * mul;
  return hash_len16(rot(e + f, 43) + rot(g, 30) + h, e + rot(f + a, 18) + g, mul);
}

// Strings longer than 64 bytes: the 64-byte-block loop.
__device__ __noinline__ uint64_t hash_long(const MemBytes& s, uint32_t len) {
  const uint64_t seed = 81;
  uint64_t x = seed;
  uint64_t y = seed * kK1 + 113;
  uint64_t z = shift_mix(y * kK2 + 113) * kK2;
  Pair v{0, 0}, w{0, 0};
  x = x * kK2 + s.fetch64(0);
  const uint32_t end = ((len - 1) / 64) * 64;
  const uint32_t last64 = end + ((len - 1) & 63) - 63;
  uint32_t pos = 0;
  do {
    x = rot(x + y + v.first + s.fetch64(pos + 8), 37) * kK1;
    y = rot(y + v.second + s.fetch64(pos + 48), 42) * kK1;
    x ^= w.second;
    y += v.first + s.fetch64(pos + 40);
    z = rot(z + w.first, 33) * kK1;
    v = weak32_at(s, pos, v.second * kK1, x + w.first);
    w = weak32_at(s, pos + 32, z + w.second, y + s.fetch64(pos + 16));
    const uint64_t t = z;
    z = x;
    x = t;
    pos += 64;
  } while (pos != end);
  const uint64_t mul = kK1 + ((z & 0xff) << 1);
  pos = last64;
  w.first += ((len - 1) & 63);
  v.first += w.first;
  w.first += v.first;
  x = rot(x + y + v.first + s.fetch64(pos + 8), 37) * mul;
  y = rot(y + v.second + s.fetch64(pos + 48), 42) * mul;
  x ^= w.second * 9;
  y += v.first * 9 + s.fetch64(pos + 40);
  z = rot(z + w.first, 33) * mul;
  v = weak32_at(s, pos, v.second * mul, x + w.first);
  w = weak32_at(s, pos + 32, z + w.second, y + s.fetch64(pos + 16));
  const uint64_t t = z;
  z = x;
  x = t;
  return hash_len16(hash_len16(v.first, w.first, mul) + shift_mix(y) * kK0 + z,
                    hash_len16(v.second, w.second, mul) + x, mul);
}

__device__ __forceinline__ uint64_t fingerprint64(const uint8_t* p, uint32_t len) {
  const MemBytes s{p};
  if (len <= 16) return hash_len_0_16(s, len);
  if (len <= 32) return hash_len_17_32(s, len);
  if (len <= 64) return hash_len_33_64(s, len);
  return hash_long(s, len);
}

// AsString for integers ("%d" / "%lld", as_string_op.cc:106-112) rendered into
// registers, then hashed. |v| has at most 19 digits, plus sign: len <= 20.
__device__ __forceinline__ uint64_t fingerprint64_of_decimal(int64_t v) {
  const bool neg = v < 0;
  uint64_t mag = neg ? (0ull - (uint64_t)v) : (uint64_t)v;
  // digits, least significant first, packed into two 64-bit words (<=10 each)
  uint32_t nd = 0;
  uint64_t d0 = 0, d1 = 0, d2 = 0;  // digit chars in reverse order
  do {
    const uint64_t q = mag / 10;
    const uint64_t ch = (uint64_t)('0' + (uint32_t)(mag - q * 10));
    const uint32_t sh = (nd & 7) * 8;
    if (nd < 8) d0 |= ch << sh;
    else if (nd < 16) d1 |= ch << sh;
    else d2 |= ch << sh;
    mag = q;
    ++nd;
  } while (mag != 0);
  if (neg) {
    const uint64_t ch = (uint64_t)'-';
    const uint32_t sh = (nd & 7) * 8;
    if (nd < 8) d0 |= ch << sh;
    else if (nd < 16) d1 |= ch << sh;
    else d2 |= ch << sh;
    ++nd;
  }
  // reverse nd bytes of (d0,d1,d2) into (w0,w1,w2)
  const RegBytes rev{d0, d1, d2};
  RegBytes s{0, 0, 0};
#pragma unroll
  for (uint32_t i = 0; i < 24; ++i) {
    if (i < nd) {
      const uint64_t ch = rev.byte(nd - 1 - i);
      const uint32_t sh = (i & 7) * 8;
      if (i < 8) s.w0 |= ch << sh;
      else if (i < 16) s.w1 |= ch << sh;
      else s.w2 |= ch << sh;
    }
  }
  if (nd <= 16) return hash_len_0_16(s, nd);
  return hash_len_17_32(s, nd);
}

// Fast path for non-negative values below 10^8 (at most 8 digits, one register).
__device__ __forceinline__ uint64_t fingerprint64_of_decimal_u32small(uint32_t v) {
  uint64_t w = 0;
  uint32_t nd = 0;
  // count digits first so characters can be placed most-significant first
  uint32_t t = v;
  do {
    t /= 10;
    ++nd;
  } while (t != 0);
  t = v;
  for (int i = (int)nd - 1; i >= 0; --i) {
    const uint32_t q = t / 10;
    w |= (uint64_t)('0' + (t - q * 10)) << (i * 8);
    t = q;
  }
  const RegBytes s{w, 0, 0};
  return hash_len_0_16(s, nd);
}

// h % n for 1 <= n < 2^31 with magic = floor(2^64 / n) (0 when n == 1):
// q = mulhi(h, magic) is floor(h / n) or one less, so one conditional subtract
// finishes; bit-exact with the `%` TF uses (string_to_hash_bucket_fast_op.h:47-55).
__device__ __forceinline__ uint64_t mod_bucket(uint64_t h, uint64_t n, uint64_t magic) {
  if (n == 1) return 0;
  const uint64_t q = __umul64hi(h, magic);
  uint64_t r = h - q * n;
  if (r >= n) r -= n;
  if (r >= n) r -= n;
  return r;
}

}  // namespace fh
}  // namespace recom
