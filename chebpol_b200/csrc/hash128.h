// hash128.h -- content fingerprint of the arrays that define an interpolant (host only).
//
// The stateless entry points re-receive the whole interpolant on every call (the R closures
// re-pass their tensor, src/chebpol.c:340-381: "no C-side state survives a call"), so the plan
// cache in api.cu has to decide from the BYTES whether it has seen this interpolant before.
// Two mechanisms, used together:
//   * a 128-bit fingerprint of every defining array (this file), computed in fixed 1 MiB
//     segments so that several host threads can share a large tensor and the value does not
//     depend on the thread count;
//   * for interpolants up to 32 MB the cache also retains a host copy and compares the bytes
//     (blobs_equal below), so a hit is exact whatever the hash does.
// For larger interpolants the fingerprint alone decides.  Each lane update h <- mix((h ^ w) * M)
// is a bijection of the lane state for a fixed word w, so two arrays that differ in any single
// word always reach different lane states; only the final 512 -> 128 bit fold can collide, with
// probability about 2^-128 per pair for inputs that are not built against the constants (the
// constants are fixed, this is not a keyed / adversarial hash).
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>
#include <omp.h>

namespace cphash {

struct H128 {
    uint64_t a = 0, b = 0;
    bool operator==(const H128 &o) const { return a == o.a && b == o.b; }
};

struct Blob {
    const void *p;
    size_t n;
};

static inline uint64_t rd64(const unsigned char *p)
{
    uint64_t v;
    memcpy(&v, p, 8);  // the int arrays (dtri, pivots, ugm_n) are only 4-byte aligned
    return v;
}
static inline uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

static const size_t kSegment = size_t(1) << 20;

// one segment (<= 1 MiB): eight independent multiply-xorshift lanes over 64-byte blocks
static inline H128 hash_segment(const unsigned char *p, size_t n, uint64_t seed)
{
    static const uint64_t M[8] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0xD6E8FEB86659FD93ull,
                                  0xA0761D6478BD642Full, 0xE7037ED1A0B428DBull, 0x8EBC6AF09C88C6E3ull, 0x589965CC75374CC3ull};
    uint64_t h[8];
    for (int j = 0; j < 8; ++j) h[j] = fmix64(seed + 0x243F6A8885A308D3ull * (uint64_t)(j + 1));
    size_t i = 0;
    for (; i + 64 <= n; i += 64)
        for (int j = 0; j < 8; ++j) {
            uint64_t v = (h[j] ^ rd64(p + i + 8 * j)) * M[j];
            h[j] = v ^ (v >> 29);
        }
    if (i < n) {
        unsigned char tail[64];
        memset(tail, 0, sizeof tail);
        memcpy(tail, p + i, n - i);
        for (int j = 0; j < 8; ++j) {
            uint64_t v = (h[j] ^ rd64(tail + 8 * j)) * M[j];
            h[j] = v ^ (v >> 29);
        }
    }
    H128 r;
    r.a = fmix64(seed ^ (uint64_t)n);
    r.b = fmix64(~seed + (uint64_t)n * M[1]);
    for (int j = 0; j < 8; ++j) {
        r.a = fmix64(r.a ^ h[j]) * M[2] + (uint64_t)j;
        r.b = fmix64(r.b + h[7 - j]) * M[3] ^ (uint64_t)(j << 8);
    }
    return r;
}

static inline void absorb(H128 &acc, const H128 &s)
{
    acc.a = fmix64(acc.a ^ s.a) * 0x9E3779B97F4A7C15ull + s.b;
    acc.b = fmix64(acc.b + s.b) * 0xC2B2AE3D27D4EB4Full ^ s.a;
}

// fingerprint of a list of arrays (order and lengths matter); `threads` host threads share the
// segments when there is enough to share
static inline H128 hash_blobs(const std::vector<Blob> &blobs, uint64_t seed, int threads)
{
    struct Seg { const unsigned char *p; size_t n; uint64_t seed; };
    std::vector<Seg> segs;
    size_t total = 0;
    for (size_t b = 0; b < blobs.size(); ++b) {
        const unsigned char *p = static_cast<const unsigned char *>(blobs[b].p);
        const size_t n = blobs[b].n;
        total += n;
        size_t off = 0, k = 0;
        do {  // an empty array still contributes a segment (its length is part of the fingerprint)
            const size_t len = n - off < kSegment ? n - off : kSegment;
            segs.push_back({p + off, len, seed + 0x100000001B3ull * (uint64_t)(b + 1) + (uint64_t)k});
            off += len;
            ++k;
        } while (off < n);
    }
    std::vector<H128> hs(segs.size());
    const int ns = (int)segs.size();
    if (threads > 1 && total >= (size_t(8) << 20)) {
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
        for (int i = 0; i < ns; ++i) hs[i] = hash_segment(segs[i].p, segs[i].n, segs[i].seed);
    } else {
        for (int i = 0; i < ns; ++i) hs[i] = hash_segment(segs[i].p, segs[i].n, segs[i].seed);
    }
    H128 acc;
    acc.a = fmix64(seed);
    acc.b = fmix64(seed ^ 0x5555555555555555ull);
    for (int i = 0; i < ns; ++i) absorb(acc, hs[i]);
    return acc;
}

static inline size_t blobs_bytes(const std::vector<Blob> &blobs)
{
    size_t t = 0;
    for (const Blob &b : blobs) t += b.n;
    return t;
}

// concatenated copy of the arrays (what the cache retains for the exact comparison)
static inline void blobs_copy(const std::vector<Blob> &blobs, std::vector<unsigned char> &dst)
{
    dst.resize(blobs_bytes(blobs));
    size_t off = 0;
    for (const Blob &b : blobs) {
        if (b.n) memcpy(dst.data() + off, b.p, b.n);
        off += b.n;
    }
}

// exact comparison of the arrays with a retained concatenated copy
static inline bool blobs_equal(const std::vector<Blob> &blobs, const std::vector<unsigned char> &kept, int threads)
{
    if (blobs_bytes(blobs) != kept.size()) return false;
    size_t off = 0;
    for (const Blob &b : blobs) {
        const unsigned char *p = static_cast<const unsigned char *>(b.p), *q = kept.data() + off;
        if (threads > 1 && b.n >= (size_t(8) << 20)) {
            const int ns = (int)((b.n + kSegment - 1) / kSegment);
            int differ = 0;
#pragma omp parallel for num_threads(threads) schedule(static) reduction(| : differ)
            for (int i = 0; i < ns; ++i) {
                const size_t lo = (size_t)i * kSegment, len = b.n - lo < kSegment ? b.n - lo : kSegment;
                differ |= memcmp(p + lo, q + lo, len) != 0;
            }
            if (differ) return false;
        } else if (b.n && memcmp(p, q, b.n) != 0) {
            return false;
        }
        off += b.n;
    }
    return true;
}

}  // namespace cphash
