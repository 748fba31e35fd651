// Wrapper that compiles the REFERENCE's own lookup3 header where it lies (REF_SRC=/root/reference/src)
// and exposes KmerHasher::getHash's arithmetic (src/Kmer.h:207-230) through a C symbol.
// Test infrastructure only: built into oracle/_ref/, used to pin oracle/kmn_oracle.c's restatement.
// The header includes libc headers inside `class Lookup3 {`, so they are pre-included here.
#include <stdio.h>
#include <time.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <sys/param.h>
#include <endian.h>
#include "lookup3.h"

extern "C" uint64_t ref_kmer_hash(const void *ptr, int length)
{
    // same seeding as KmerHasher::getHash: hash=0xDEADBEEF, pc=low word, pb=high word
    uint64_t hash = 0xDEADBEEF;
    uint32_t pc, pb;
    memcpy(&pc, &hash, 4);
    memcpy(&pb, ((char *)&hash) + 4, 4);
    Lookup3::hashlittle2(ptr, (size_t)length, &pc, &pb);
    return (uint64_t)pc | ((uint64_t)pb << 32);
}
