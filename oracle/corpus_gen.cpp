/*
 * corpus_gen.cpp -- TEST / BENCH INFRASTRUCTURE.  Host build of the deterministic synthetic manifest generator
 * (operator-builder_b200/csrc/obm_corpus.h, the same source the device generator compiles) as a library of its
 * own, so that bench.py's reference arm and CPU baseline leg get their corpus without mapping libobmarkers.so:
 * that process then holds oracle code only.  No lexing here.
 */
#include <cstdint>
#include "../operator-builder_b200/csrc/obm_corpus.h"

extern "C" int obc_generate_corpus_host(uint8_t *bytes, uint64_t *doc_off, uint32_t ndocs, uint32_t doc_bytes, uint64_t first_doc, int flavour) {
    if (!bytes) return -4;
    for (uint32_t d = 0; d < ndocs; d++) {
        if (doc_off) doc_off[d] = (uint64_t)d * doc_bytes;
        obmc::generate_doc(bytes + (uint64_t)d * doc_bytes, doc_bytes, first_doc + d, flavour);
    }
    if (doc_off) doc_off[ndocs] = (uint64_t)ndocs * doc_bytes;
    return 0;
}
