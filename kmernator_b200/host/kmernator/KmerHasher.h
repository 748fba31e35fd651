// KmerHasher.h -- host-side KmerHasher::getHash (src/Kmer.h:207-230): Bob Jenkins' lookup3 hashlittle2 (public domain,
// src/lookup3.h:470-644 in the reference) over the (k+3)/4 key bytes with pc = 0xDEADBEEF, pb = 0, result c | b << 32.
// Needed on the host only for what depends on the reference's BUCKET ORDER: the --save-kmer-mmap file layout
// (bucket = hash & mask, src/Kmer.h:2329-2333) and the owner rank of a key (src/Kmer.h:2284-2295).  Restated for byte
// strings of any length from the published algorithm (little-endian byte reads, no alignment assumptions).
#ifndef KMERNATOR_HOST_KMERHASHER_H
#define KMERNATOR_HOST_KMERHASHER_H

#include <cstddef>
#include <cstdint>

namespace kmn_host {

inline uint32_t rot32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

inline uint64_t kmerHash(const uint8_t *key, size_t length)
{
    uint32_t a, b, c;
    a = b = c = 0xdeadbeefu + (uint32_t)length + 0xDEADBEEFu;
    // c += pb (0)
    auto le = [&](size_t off, size_t n) {                                // up to 4 bytes, little endian, zero padded
        uint32_t v = 0;
        for (size_t i = 0; i < n && i < 4; ++i) v |= (uint32_t)key[off + i] << (8 * i);
        return v;
    };
    size_t off = 0;
    while (length > 12) {
        a += le(off, 4); b += le(off + 4, 4); c += le(off + 8, 4);
        a -= c; a ^= rot32(c, 4);  c += b;
        b -= a; b ^= rot32(a, 6);  a += c;
        c -= b; c ^= rot32(b, 8);  b += a;
        a -= c; a ^= rot32(c, 16); c += b;
        b -= a; b ^= rot32(a, 19); a += c;
        c -= b; c ^= rot32(b, 4);  b += a;
        length -= 12; off += 12;
    }
    if (length == 0) return (uint64_t)c | ((uint64_t)b << 32);
    a += le(off, length);
    if (length > 4) b += le(off + 4, length - 4);
    if (length > 8) c += le(off + 8, length - 8);
    c ^= b; c -= rot32(b, 14);
    a ^= c; a -= rot32(c, 11);
    b ^= a; b -= rot32(a, 25);
    c ^= b; c -= rot32(b, 16);
    a ^= c; a -= rot32(c, 4);
    b ^= a; b -= rot32(a, 14);
    c ^= b; c -= rot32(b, 24);
    return (uint64_t)c | ((uint64_t)b << 32);
}

// getDistributedThreadId: ((hash >> 24) & 0x7ffff) % numRanks           src/Kmer.h:187-188,2284-2295
inline uint32_t kmerOwner(uint64_t hash, uint32_t nranks) { return (uint32_t)((hash >> 24) & 0x7ffffu) % nranks; }

}  // namespace kmn_host

#endif
