/*
 * obm_tile.h -- the tile fast path's per-thread phase logic, written as host/device functions over
 * a shared-memory image (`obmt::Smem`) so that tests/hostsim can replay a CTA phase by phase on the
 * CPU.  The CUDA kernel (obm_fast.cuh) supplies the data movement (TMA bulk copy into Smem::data),
 * the barriers between phases, the warp-shuffle scans and the cross-CTA look-back.
 *
 * One CTA owns the documents that START inside a TILE-byte range of the packed batch.  Documents of
 * at most MAXDOC bytes are staged whole in shared memory and lexed line-parallel:
 *
 *   P2 classify     every 32-byte word -> newline bitmap, special bitmap ({# ' + /}: everything that
 *                   can start a comment or a marker), non-ASCII flag
 *   P3 doc prep     a virtual newline in front of every document start; per-document non-ASCII flag
 *   P4 line scan    newline prefix counts (line numbers) + "owners": lines whose first special byte
 *                   exists, found by looking forward from every (virtual) newline
 *   P5 owners       one thread per owner runs obm::Lexer in LINE mode from the line's first special
 *                   byte (reference semantics: a line starts in state `lex`), skipping
 *                   dull bytes through the bitmaps
 *   P6 resolve      documents whose lines interact (multi-line literal, fatal error) or that contain
 *                   non-ASCII bytes are re-lexed sequentially by one thread (exact path, same core);
 *                   tuple counts -> offsets
 *   P7 fill         owners re-run writing tuples at their final global positions
 *
 * The tuple stream is identical to the exact path's by construction (same obm::Lexer, canonical
 * LINE/PART rules); tests compare the two streams tuple for tuple.
 */
#ifndef OBM_TILE_H
#define OBM_TILE_H

#include <stdint.h>
#include "obm_core.h"

namespace obmt {

constexpr uint32_t NT = 256;        /* threads per CTA */
constexpr uint32_t TILE = 16384;    /* a CTA owns the documents starting in [t*TILE, (t+1)*TILE) */
constexpr uint32_t MAXDOC = 16368;  /* documents up to this size take the tile path */
constexpr uint32_t NW = 1024;       /* 32-byte words staged per sub-batch (32 KiB) */
constexpr uint32_t WPT = NW / NT;   /* words per thread in the line scan (4) */
constexpr uint32_t DMAX = 64;       /* documents per sub-batch */
constexpr uint32_t QMAX = 1024;     /* owners per sub-batch */

enum : uint32_t { DF_NONASCII = 1, DF_INTERACT = 2, DF_QOVERFLOW = 4 };

struct Smem {
    alignas(16) uint8_t data[NW * 32];
    alignas(16) uint32_t nlw[NW];   /* bit i of word w: byte 32w+i is '\n' (or precedes a document start) */
    alignas(16) uint32_t spw[NW];   /* bit i of word w: byte 32w+i is one of # ' + / */
    uint16_t nlpre[NW];             /* number of nlw bits in words [0, w) */
    uint32_t naw[NW / 32];          /* bit w%32 of naw[w/32]: word w holds a byte >= 0x80 */
    uint32_t owner[QMAX];           /* first special position | line start << 16 (buffer-relative) */
    uint32_t ocnt[QMAX + 1];        /* tuple count of the owner; after the scan: exclusive prefix E[] */
    uint8_t odoc[QMAX];             /* document (index in the sub-batch) of the owner */
    uint32_t dstart[DMAX + 1];      /* document start positions, buffer-relative; [nd] = end */
    uint32_t dflag[DMAX];
    uint32_t dcnt[DMAX + 1];        /* tuples per document; after the scan: exclusive offsets */
    uint32_t dfirst[DMAX + 1];      /* first owner of each document */
    uint32_t scan_tmp[NT / 32 + 1];
    uint32_t n_owners;
    uint32_t nd;
    uint32_t lo_pos, hi_pos;        /* valid byte range of `data` */
};

/* ---- P2: classification ---------------------------------------------------------------------- */
/* 4-bit mask: bit b set iff byte b of t is zero (t's bytes must be < 0x80) */
OBM_HD uint32_t zero_bytes4(uint32_t t) {
    uint32_t ne = ((t + 0x7F7F7F7Fu) >> 7) & 0x01010101u; /* 1 per non-zero byte */
    uint32_t eq = ne ^ 0x01010101u;
    return (eq * 0x00204081u >> 21) & 0xFu;
}

OBM_HD void classify_word(Smem &S, uint32_t wi) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(S.data) + wi * 8;
    uint32_t nl = 0, sp = 0, hi = 0;
#if defined(__CUDA_ARCH__)
    const uint4 a = reinterpret_cast<const uint4 *>(src)[0], b = reinterpret_cast<const uint4 *>(src)[1];
    const uint32_t x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#else
    const uint32_t *x = src;
#endif
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t v = x[k];
        hi |= v;
        nl |= zero_bytes4(v ^ 0x0A0A0A0Au) << (4 * k);
        sp |= zero_bytes4((v & 0xF3F3F3F3u) ^ 0x23232323u) << (4 * k);
    }
    /* keep only bytes inside [lo_pos, hi_pos) */
    uint32_t w0 = wi * 32;
    uint32_t keep = 0xFFFFFFFFu;
    if (S.lo_pos > w0) keep &= (S.lo_pos - w0 >= 32) ? 0u : (0xFFFFFFFFu << (S.lo_pos - w0));
    if (S.hi_pos < w0 + 32) keep &= (S.hi_pos <= w0) ? 0u : (0xFFFFFFFFu >> (w0 + 32 - S.hi_pos));
    S.nlw[wi] = nl & keep;
    S.spw[wi] = sp & keep;
    bool na = (hi & 0x80808080u) != 0 && keep != 0;
#if defined(__CUDA_ARCH__)
    uint32_t bal = __ballot_sync(0xFFFFFFFFu, na); /* callers keep warps converged and wi = base + lane */
    if ((wi & 31) == 0) S.naw[wi >> 5] = bal;
#else
    if ((wi & 31) == 0) S.naw[wi >> 5] = 0;
    if (na) S.naw[wi >> 5] |= 1u << (wi & 31);
#endif
}

OBM_HD void atomic_or_u32(uint32_t *p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}

/* ---- P3: per-document preparation (thread d < nd) ---------------------------------------------- */
OBM_HD void doc_prep(Smem &S, uint32_t d) {
    uint32_t q = S.dstart[d], e = S.dstart[d + 1];
    if (q > S.lo_pos) atomic_or_u32(&S.nlw[(q - 1) >> 5], 1u << ((q - 1) & 31)); /* virtual newline before the document */
    uint32_t flag = 0;
    if (e > q) {
        uint32_t w0 = q >> 5, w1 = (e - 1) >> 5; /* words touched, boundary words included (conservative) */
        for (uint32_t g = w0 >> 5; g <= (w1 >> 5); g++) {
            uint32_t m = S.naw[g];
            uint32_t lo = g * 32, hiw = lo + 31;
            if (w0 > lo) m &= 0xFFFFFFFFu << (w0 - lo);
            if (w1 < hiw) m &= 0xFFFFFFFFu >> (hiw - w1);
            if (m) { flag = DF_NONASCII; break; }
        }
    }
    S.dflag[d] = flag;
}

/* ---- P4: line scan (thread t owns words [WPT*t, WPT*t+WPT)) -------------------------------------- */
/* first position >= from whose (sp|nl) bit is set, or hi_pos if none */
OBM_HD uint32_t next_event(const Smem &S, uint32_t from) {
    if (from >= S.hi_pos) return S.hi_pos;
    uint32_t w = from >> 5;
    uint32_t m = (S.spw[w] | S.nlw[w]) & (0xFFFFFFFFu << (from & 31));
    const uint32_t wend = (S.hi_pos + 31) >> 5;
    while (m == 0) {
        if (++w >= wend) return S.hi_pos;
        m = S.spw[w] | S.nlw[w];
    }
#if defined(__CUDA_ARCH__)
    uint32_t pos = w * 32 + (uint32_t)(__ffs((int)m) - 1);
#else
    uint32_t pos = w * 32 + (uint32_t)__builtin_ctz(m);
#endif
    return pos < S.hi_pos ? pos : S.hi_pos;
}

/* Visits the owners whose line starts right after a newline bit in this thread's words (plus the
 * line starting at lo_pos).  f(first_special_pos, line_start) is called in position order.
 * Returns the number of nlw bits in the thread's words. */
template <class F>
OBM_HD uint32_t line_scan(const Smem &S, uint32_t t, F &&f) {
    uint32_t nls = 0;
#pragma unroll
    for (uint32_t j = 0; j < WPT; j++) {
        uint32_t w = t * WPT + j;
        uint32_t nl = S.nlw[w];
        if (S.lo_pos < S.hi_pos && (S.lo_pos >> 5) == w) { /* the buffer's first line has no newline before it */
            uint32_t ev = next_event(S, S.lo_pos);
            if (ev < S.hi_pos && ((S.spw[ev >> 5] >> (ev & 31)) & 1u)) f(ev, S.lo_pos);
        }
#if defined(__CUDA_ARCH__)
        nls += (uint32_t)__popc(nl);
#else
        nls += (uint32_t)__builtin_popcount(nl);
#endif
        while (nl) {
#if defined(__CUDA_ARCH__)
            uint32_t b = (uint32_t)(__ffs((int)nl) - 1);
#else
            uint32_t b = (uint32_t)__builtin_ctz(nl);
#endif
            nl &= nl - 1;
            uint32_t start = w * 32 + b + 1;
            uint32_t ev = next_event(S, start);
            if (ev < S.hi_pos && ((S.spw[ev >> 5] >> (ev & 31)) & 1u)) f(ev, start);
        }
    }
    return nls;
}

/* number of nlw bits at positions < q */
OBM_HD uint32_t nl_before(const Smem &S, uint32_t q) {
    uint32_t w = q >> 5;
    if (w >= NW) return (uint32_t)S.nlpre[NW - 1] +
#if defined(__CUDA_ARCH__)
        (uint32_t)__popc(S.nlw[NW - 1]);
#else
        (uint32_t)__builtin_popcount(S.nlw[NW - 1]);
#endif
    uint32_t m = S.nlw[w] & ((1u << (q & 31)) - 1u);
#if defined(__CUDA_ARCH__)
    return (uint32_t)S.nlpre[w] + (uint32_t)__popc(m);
#else
    return (uint32_t)S.nlpre[w] + (uint32_t)__builtin_popcount(m);
#endif
}

/* ---- P5 / P7: owners ------------------------------------------------------------------------- */
struct BitmapAccel {
    const Smem *S; uint32_t dpos, dend; /* document start / end, buffer-relative */
    OBM_HD uint32_t next_interesting(uint32_t p) const {
        uint32_t q = next_event(*S, dpos + p);
        if (q > dend) q = dend;
        return q - dpos;
    }
};

/* document (index in the sub-batch) containing buffer position `pos`: the last d with dstart[d] <= pos */
OBM_HD uint32_t doc_of(const Smem &S, uint32_t pos) {
    uint32_t lo = 0, hi = S.nd; /* invariant: dstart[lo] <= pos; answer in [lo, hi) */
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (S.dstart[mid] <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

struct OwnerResult { uint32_t tuples, markers, lexemes; bool interact; };

/* Runs owner `o` with the given sink.  The sink sees document-relative offsets. */
template <class Sink>
OBM_HD OwnerResult owner_run(const Smem &S, const obm::Tables &T, uint32_t o, uint32_t d, Sink &sink) {
    uint32_t rec = S.owner[o];
    uint32_t first = rec & 0xFFFFu, ls = rec >> 16;
    uint32_t dpos = S.dstart[d], dend = S.dstart[d + 1];
    uint32_t line = 1 + nl_before(S, ls) - nl_before(S, dpos);
    BitmapAccel acc{&S, dpos, dend};
    obm::Lexer<Sink, BitmapAccel> lx(T, S.data + dpos, dend - dpos, sink, first - dpos, line, ls - dpos,
                                     !(line == 1 && ls == dpos), acc);
    int st = lx.template run<true>();
    OwnerResult r;
    r.tuples = (uint32_t)sink.n_tuples; r.markers = sink.n_markers; r.lexemes = sink.n_lexemes;
    uint32_t end_line = lx.line_p - (st == obm::RUN_LINE_END ? 1u : 0u);
    r.interact = (st == obm::RUN_FATAL) || (end_line != line);
    return r;
}

/* P5 body for owner o: count pass */
OBM_HD void owner_count(Smem &S, const obm::Tables &T, uint32_t o) {
    uint32_t d = doc_of(S, S.owner[o] >> 16);
    S.odoc[o] = (uint8_t)d;
    if (S.dflag[d]) { S.ocnt[o] = 0; return; }
    obm::CountSink sink;
    OwnerResult r = owner_run(S, T, o, d, sink);
    S.ocnt[o] = r.tuples;
    if (r.interact) atomic_or_u32(&S.dflag[d], DF_INTERACT);
}

/* first owner whose line start is >= pos */
OBM_HD uint32_t first_owner_at(const Smem &S, uint32_t pos) {
    uint32_t lo = 0, hi = S.n_owners;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((S.owner[mid] >> 16) < pos) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* P6 body for document d (after the owner-count scan turned ocnt[] into E[]): tuples of the document */
OBM_HD void doc_count(Smem &S, const obm::Tables &T, uint32_t d) {
    uint32_t dpos = S.dstart[d], dend = S.dstart[d + 1];
    if (S.dflag[d]) { /* exact path, sequential, same core */
        obm::CountSink sink;
        obm::Lexer<obm::CountSink> lx(T, S.data + dpos, dend - dpos, sink);
        lx.template run<false>();
        S.dcnt[d] = (uint32_t)sink.n_tuples;
    } else {
        S.dcnt[d] = S.ocnt[S.dfirst[d + 1]] - S.ocnt[S.dfirst[d]] + 1; /* + EOF */
    }
}

struct FillStats { uint32_t markers, lexemes, exact_docs, fatal_docs; };

/* P7 body for owner o: writes its tuples; `doc_base[d]` = global tuple index of the document's first tuple */
OBM_HD void owner_fill(const Smem &S, const obm::Tables &T, uint32_t o, obm_tuple *out, uint64_t out_cap,
                       uint64_t batch_base, FillStats &fs) {
    uint32_t d = S.odoc[o];
    if (S.dflag[d]) return;
    uint64_t at = batch_base + S.dcnt[d] + (S.ocnt[o] - S.ocnt[S.dfirst[d]]);
    obm::WriteSink sink(out + at, at < out_cap ? out_cap - at : 0);
    OwnerResult r = owner_run(S, T, o, d, sink);
    fs.markers += r.markers; fs.lexemes += r.lexemes;
}

/* P7 body for document d: EOF tuple of a regular document, or the whole irregular document */
OBM_HD void doc_fill(const Smem &S, const obm::Tables &T, uint32_t d, obm_tuple *out, uint64_t out_cap,
                     uint64_t batch_base, FillStats &fs) {
    uint32_t dpos = S.dstart[d], dend = S.dstart[d + 1];
    uint64_t at = batch_base + S.dcnt[d];
    if (S.dflag[d]) {
        obm::WriteSink sink(out + at, at < out_cap ? out_cap - at : 0);
        obm::Lexer<obm::WriteSink> lx(T, S.data + dpos, dend - dpos, sink);
        int st = lx.template run<false>();
        fs.markers += sink.n_markers; fs.lexemes += sink.n_lexemes; fs.exact_docs += 1; fs.fatal_docs += (st == obm::RUN_FATAL);
    } else {
        uint64_t eof_at = batch_base + S.dcnt[d + 1] - 1;
        if (eof_at < out_cap) out[eof_at] = OBM_TUPLE(OBM_K_EOF, dend - dpos, 0);
        fs.lexemes += 1;
    }
}

} /* namespace obmt */
#endif
