/*
 * hostsim.cpp -- TEST INFRASTRUCTURE.  Compiles the product's lexer core (obm_core.h, the same
 * source the CUDA kernels instantiate) for the host so that `pytest -m "not gpu"` can check the
 * tuple stream + decoder logic against the oracle without a GPU.  Never loaded by the product
 * package; libobmarkers.so has no host lexing entry point.
 */
#include <cstdint>
#include <cstring>
#include "../../operator-builder_b200/csrc/go_unicode_tables.h"
#include "../../operator-builder_b200/csrc/obm_core.h"

static const char F64[] = GO_F64_OVERFLOW_DIGITS;
static const obm::Tables TBL = { GO_LETTER_RANGES, GO_LETTER_RANGES_N, GO_NUMBER_RANGES, GO_NUMBER_RANGES_N, F64 };

extern "C" {
/* whole-document exact path; returns the tuple count (writes at most cap) */
uint64_t hs_lex_doc(const uint8_t *doc, uint32_t n, obm_tuple *out, uint64_t cap, uint32_t *n_markers, uint32_t *n_lexemes) {
    obm::WriteSink sink(out, cap);
    obm::Lexer<obm::WriteSink> lx(TBL, doc, n, sink);
    lx.run<false>();
    if (n_markers) *n_markers = sink.n_markers;
    if (n_lexemes) *n_lexemes = sink.n_lexemes;
    return sink.n_tuples;
}
/* per-line composition, the way the fast kernel assembles a document: one LINE_MODE lexer per
 * physical line that has not been swallowed by an earlier owner, EOF appended unless fatal. */
uint64_t hs_lex_doc_by_lines(const uint8_t *doc, uint32_t n, obm_tuple *out, uint64_t cap) {
    obm::WriteSink sink(out, cap);
    uint32_t pos = 0, line = 1;
    bool fatal = false;
    while (pos < n) {
        uint32_t le = pos;
        while (le < n && doc[le] != '\n') le++;
        obm::Lexer<obm::WriteSink> lx(TBL, doc, n, sink, pos, line, !(line == 1 && pos == 0));
        int st = lx.run<true>(le);
        if (st == obm::RUN_FATAL) { fatal = true; break; }
        if (st == obm::RUN_EOF) break;
        pos = lx.p; line = lx.line_p;
    }
    if (!fatal) sink.put(OBM_K_EOF, n, 0);
    return sink.n_tuples;
}
int hs_parse_float_err(const uint8_t *s, uint32_t n) { return obm::parse_float_err(TBL, s, n); }
int hs_atoi_err(const uint8_t *s, uint32_t n) { return obm::atoi_err(s, n); }
}
