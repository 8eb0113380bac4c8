/*
 * obm_parse_dev.h -- the lexer's only consumer, internal/markers/parser, ON THE DEVICE: one
 * thread per document walks the document's tuples the way parser/state.go:13-175 walks lexemes and emits compact
 * Result records instead of Go structs:
 *   registry lookup            definition.go:13-21 (marker name = scopeBuffer minus the trailing ':')
 *   known-argument filter      state.go:79-93 (LookupArgument)
 *   value typing               state.go:95-153 (ParseBool, Atoi class, ParseFloat(., 32) range)
 *   MarkerText                 position.go:18 + emit.go:8-24, as a span of the document (+ the synthetic "\n")
 *   error results              error.go:8-22 (they end the document's parse, parser.go:63-73)
 * The same source is the host mirror's fast path (obm_parse.cpp: obm_results_format_doc) and is checked against
 * obm_parse_doc / oracle/parser_oracle.py in tests/test_parser_row.py.
 *
 * Scope of the device walk: documents whose tuple stream holds only plain lexemes (kinds 1..20) and LINE tuples.
 * Anything that makes a Value differ from an input slice or needs the decoder's text formatting -- PART / FLUSH /
 * DRIFT / LINEHI pseudo-tuples, in-band warnings and lexer errors -- marks the whole document OBM_R_HOST: one record,
 * and the host runs obm_parse_doc on that document's tuples (exact for everything).
 */
#ifndef OBM_PARSE_DEV_H
#define OBM_PARSE_DEV_H

#include "obm_core.h"

namespace obmr {

struct DevRegistry { /* names and argument names back to back in text[] */
    uint32_t n; uint32_t name_off[9]; uint32_t arg_first[9]; uint32_t arg_off[65]; uint8_t text[1024];
};

OBM_HD bool slice_eq(const uint8_t *doc, uint32_t off, uint32_t len, const uint8_t *t, uint32_t tlen) {
    if (len != tlen) return false;
    uint32_t diff = 0;
    for (uint32_t k = 0; k < len; k++) diff |= (uint32_t)doc[off + k] ^ (uint32_t)t[k]; /* independent loads: they pipeline */
    return diff == 0;
}
OBM_HD int lookup_marker(const DevRegistry &R, const uint8_t *doc, uint32_t off, uint32_t len) {
    for (uint32_t r = 0; r < R.n; r++) if (slice_eq(doc, off, len, R.text + R.name_off[r], R.name_off[r + 1] - R.name_off[r])) return (int)r;
    return -1;
}
OBM_HD bool lookup_arg(const DevRegistry &R, uint32_t def, const uint8_t *doc, uint32_t off, uint32_t len) {
    for (uint32_t a = R.arg_first[def]; a < R.arg_first[def + 1]; a++)
        if (slice_eq(doc, off, len, R.text + R.arg_off[a], R.arg_off[a + 1] - R.arg_off[a])) return true;
    return false;
}
/* strconv.ParseBool accepts exactly: 1 t T TRUE true True 0 f F FALSE false False */
OBM_HD bool parse_bool_ok(const uint8_t *v, uint32_t n) {
    if (n == 1) return v[0] == '1' || v[0] == 't' || v[0] == 'T' || v[0] == '0' || v[0] == 'f' || v[0] == 'F';
    if (n == 4) return (v[0] == 't' && v[1] == 'r' && v[2] == 'u' && v[3] == 'e') || (v[0] == 'T' && v[1] == 'R' && v[2] == 'U' && v[3] == 'E') ||
                       (v[0] == 'T' && v[1] == 'r' && v[2] == 'u' && v[3] == 'e');
    if (n == 5) return (v[0] == 'f' && v[1] == 'a' && v[2] == 'l' && v[3] == 's' && v[4] == 'e') || (v[0] == 'F' && v[1] == 'A' && v[2] == 'L' && v[3] == 'S' && v[4] == 'E') ||
                       (v[0] == 'F' && v[1] == 'a' && v[2] == 'l' && v[3] == 's' && v[4] == 'e');
    return false;
}
/* does the (lexer-validated) decimal literal overflow float32?  |value| >= 2^128 - 2^103 = 3.40282356779733661637539395458142568448e38 */
OBM_HD bool float32_overflows(const uint8_t *v, uint32_t n) {
    const char *T = "340282356779733661637539395458142568448";
    uint32_t i = 0;
    if (i < n && (v[i] == '+' || v[i] == '-')) i++;
    const uint32_t dbeg = i;
    long long dp = 0; bool dot = false, started = false; uint32_t nd = 0;
    for (; i < n; i++) {
        const uint32_t c = v[i];
        if (c == '.') { dot = true; continue; }
        if (c < '0' || c > '9') break;
        if (!started && c == '0') { if (dot) dp--; continue; }
        started = true; nd++;
        if (!dot) dp++;
    }
    const uint32_t dend = i;
    if (!started) return false;
    if (i < n && (v[i] == 'e' || v[i] == 'E')) {
        i++; int sg = 1; long long e = 0;
        if (i < n && (v[i] == '+' || v[i] == '-')) { if (v[i] == '-') sg = -1; i++; }
        for (; i < n && v[i] >= '0' && v[i] <= '9'; i++) if (e < 10000) e = e * 10 + (v[i] - '0');
        dp += sg * e;
    }
    if (dp > 39) return true;
    if (dp < 39) return false;
    uint32_t k = 0; bool st = false;
    for (uint32_t j = dbeg; j < dend && k < 39; j++) {
        const uint32_t c = v[j];
        if (c == '.') continue;
        if (!st) { if (c == '0') continue; st = true; }
        if (c > (uint32_t)T[k]) return true;
        if (c < (uint32_t)T[k]) return false;
        k++;
    }
    for (; k < 39; k++) { if ('0' > T[k]) return true; if ('0' < T[k]) return false; }
    (void)nd;
    return true;
}

/* sink: counts always, writes when the arrays are given */
struct Sink {
    obm_result *res; uint64_t res_cap; obm_arg *args; uint64_t arg_cap;
    uint64_t res_at, arg_at; /* next free index (batch-global) */
    uint32_t nres, nargs;
    OBM_HD void arg(uint32_t name_off, uint32_t name_len, uint32_t kind, uint32_t val_off, uint32_t val_len, uint32_t flags) {
        if (args && arg_at + nargs < arg_cap) {
            obm_arg a; a.name_off = name_off; a.val_off = val_off; a.val_len = val_len; a.name_len = (uint16_t)name_len; a.kind = (uint8_t)kind; a.flags = (uint8_t)flags;
            args[arg_at + nargs] = a;
        }
        nargs++;
    }
    OBM_HD void result(uint32_t doc, uint32_t tuple, uint32_t text_off, uint32_t text_len, uint32_t reg_id, uint32_t n_args, uint64_t arg_base, uint32_t flags, uint32_t aux) {
        if (res && res_at + nres < res_cap) {
            obm_result r; r.doc = doc; r.tuple = tuple; r.text_off = text_off; r.text_len = text_len; r.reg_id = (uint16_t)reg_id; r.nargs = (uint16_t)n_args;
            r.arg_base = (uint32_t)arg_base; r.flags = flags; r.aux = aux;
            res[res_at + nres] = r;
        }
        nres++;
    }
};

/* One document.  tuples: the document's slice of the stream.  doc_id: what goes into the records. */
OBM_HD void parse_doc(const DevRegistry &R, const uint8_t *doc, const obm_tuple *t, uint32_t nt, uint32_t doc_id, Sink &S) {
    /* pass 0: anything the device walk does not model? */
    for (uint32_t i = 0; i < nt; i++) {
        const uint32_t k = OBM_TUPLE_KIND(t[i]);
        if (k > OBM_K_EOF && k != OBM_K_LINE) { S.result(doc_id, i, 0, 0, 0xFFFFu, 0, 0, OBM_R_HOST, k); return; }
    }
    enum St { S_START, S_PARSE, S_MARKER_START, S_SCOPE, S_SEPARATOR, S_ARG, S_MORE, S_STOP };
    uint32_t i = 0;                                  /* next lexeme (LINE tuples are skipped on the way) */
    uint32_t line = 1, base = 0;                     /* position basis of the lexeme at i */
    uint32_t sc_off = 0, sc_len = 0; bool sc_any = false, sc_nl = false, sc_broken = false; /* scopeBuffer as a span (+ "\n") */
    uint32_t cur_tuple = 0, cur_line = 0, cur_col = 0; /* parser.current */
    int def = -1; uint32_t marker_tuple = 0, n_args = 0; uint64_t arg_base = 0;
    auto skip_line = [&]() { while (i < nt && OBM_TUPLE_KIND(t[i]) == OBM_K_LINE) { base = OBM_TUPLE_OFF(t[i]); line = OBM_TUPLE_LEN(t[i]); i++; } };
    auto kind_at = [&]() -> uint32_t { skip_line(); return i < nt ? OBM_TUPLE_KIND(t[i]) : 0u /* closed channel: zero Lexeme */; };
    auto scope_clear = [&]() { sc_any = false; sc_nl = false; sc_len = 0; sc_broken = false; };
    auto next = [&]() { /* position.go:7-20: consume the lexeme, append its Value to scopeBuffer */
        skip_line();
        if (i >= nt) { cur_tuple = nt; cur_line = cur_col = 0; return; }
        const obm_tuple tu = t[i];
        const uint32_t k = OBM_TUPLE_KIND(tu), off = OBM_TUPLE_OFF(tu), len = OBM_TUPLE_LEN(tu);
        cur_tuple = i;
        const bool synthetic = (k == OBM_K_SYNTHETIC_BOOL || k == OBM_K_MARKER_END || k == OBM_K_EOF);
        if (synthetic) { cur_line = cur_col = 0; } else { cur_line = line; cur_col = off - base + 1; }
        if (k == OBM_K_MARKER_END) { if (sc_nl) sc_broken = true; sc_nl = true; if (!sc_any) { sc_any = true; sc_off = off; sc_len = 0; } }
        else if (k == OBM_K_SYNTHETIC_BOOL) sc_broken = true;   /* "true" is not input text (only reachable through the catch-all of parse) */
        else if (len) {
            if (!sc_any) { sc_any = true; sc_off = off; sc_len = len; }
            else if (sc_nl || off != sc_off + sc_len) sc_broken = true;
            else sc_len += len;
        }
        i++;
    };
    auto discard = [&]() { skip_line(); if (i < nt) i++; };
    uint32_t st = S_START;
    while (st != S_STOP) {
        const uint32_t k = kind_at();
        switch (st) {
        case S_START: case S_PARSE: /* state.go:13-46 */
            if (k == OBM_K_COMMENT && i < nt) { discard(); st = S_PARSE; }
            else if (k == OBM_K_MARKER_START && i < nt) { marker_tuple = i; next(); st = S_MARKER_START; }
            else if (k == OBM_K_EOF && i < nt) { next(); st = S_STOP; }
            else if (i >= nt) { st = S_STOP; /* closed channel: the zero lexeme is an Error lexeme with an empty value -- cannot happen: EOF or a fatal error ends every stream */ }
            else if (st == S_START) st = S_PARSE;
            else { next(); scope_clear(); st = S_PARSE; }
            break;
        case S_MARKER_START: if (k == OBM_K_SCOPE && i < nt) { next(); st = S_SCOPE; } else st = S_PARSE; break;
        case S_SCOPE: if (k == OBM_K_SEPARATOR && i < nt) { next(); st = S_SEPARATOR; } else st = S_PARSE; break;
        case S_SEPARATOR: /* state.go:64-77 */
            if (k == OBM_K_SCOPE && i < nt) { next(); st = S_SCOPE; break; }
            if (k == OBM_K_ARG && i < nt && sc_any && (sc_len || sc_nl)) {
                if (sc_broken || sc_nl) { S.nres = 0; S.nargs = 0; S.result(doc_id, i, 0, 0, 0xFFFFu, 0, 0, OBM_R_HOST, 0); return; }
                def = lookup_marker(R, doc, sc_off, sc_len - 1);
                if (def >= 0) { n_args = 0; arg_base = S.arg_at + S.nargs; st = S_ARG; break; }
            }
            scope_clear(); def = -1; st = S_PARSE;
            break;
        case S_ARG: { /* state.go:79-93 + parseArgValue :95-153 */
            if (!(k == OBM_K_ARG && i < nt)) { st = S_PARSE; break; }
            const uint32_t a_off = OBM_TUPLE_OFF(t[i]), a_len = OBM_TUPLE_LEN(t[i]);
            next();
            if (!lookup_arg(R, (uint32_t)def, doc, a_off, a_len)) { st = S_PARSE; break; }
            if (kind_at() == OBM_K_ARG_ASSIGNMENT && i < nt) next();
            if (kind_at() == OBM_K_QUOTE && i < nt) next();
            const uint32_t vk = kind_at();
            const uint32_t v_off = i < nt ? OBM_TUPLE_OFF(t[i]) : 0u, v_len = i < nt ? OBM_TUPLE_LEN(t[i]) : 0u;
            if (i >= nt) { st = S_PARSE; break; }
            if (vk == OBM_K_SYNTHETIC_BOOL) { S.arg(a_off, a_len, 0, v_off, 0, OBM_A_SYNTHETIC_TRUE); n_args++; discard(); }
            else if (vk == OBM_K_BOOL_LITERAL) {
                next();
                if (!parse_bool_ok(doc + v_off, v_len)) {
                    S.nargs -= n_args; /* an error result replaces the marker's arguments */
                    const uint64_t ab = S.arg_at + S.nargs;
                    S.arg(cur_line, 0, 0, v_off, v_len, 0);
                    S.result(doc_id, cur_tuple, sc_off, sc_len, (uint32_t)def, 1, ab, OBM_R_ERR_PARSEBOOL | (sc_broken ? OBM_R_HOST : 0u), cur_col);
                    if (sc_broken) { S.nres = 0; S.nargs = 0; S.result(doc_id, i, 0, 0, 0xFFFFu, 0, 0, OBM_R_HOST, 0); }
                    return;
                }
                S.arg(a_off, a_len, 0, v_off, v_len, 0); n_args++;
            } else if (vk == OBM_K_INTEGER_LITERAL) { next(); S.arg(a_off, a_len, 1, v_off, v_len, 0); n_args++; }
            else if (vk == OBM_K_FLOAT_LITERAL) {
                next();
                if (float32_overflows(doc + v_off, v_len)) {
                    S.nargs -= n_args;
                    const uint64_t ab = S.arg_at + S.nargs;
                    S.arg(cur_line, 0, 2, v_off, v_len, 0);
                    S.result(doc_id, cur_tuple, sc_off, sc_len, (uint32_t)def, 1, ab, OBM_R_ERR_FLOAT32 | (sc_broken ? OBM_R_HOST : 0u), cur_col);
                    if (sc_broken) { S.nres = 0; S.nargs = 0; S.result(doc_id, i, 0, 0, 0xFFFFu, 0, 0, OBM_R_HOST, 0); }
                    return;
                }
                S.arg(a_off, a_len, 2, v_off, v_len, 0); n_args++;
            } else if (vk == OBM_K_STRING_LITERAL) {
                next(); S.arg(a_off, a_len, 3, v_off, v_len, 0); n_args++;
                if (kind_at() == OBM_K_QUOTE && i < nt) next();
            } else { st = S_PARSE; break; }
            st = S_MORE;
            break;
        }
        case S_MORE: /* state.go:155-169 */
            if (k == OBM_K_ARG_DELIMITER && i < nt) { next(); st = S_ARG; }
            else if (k == OBM_K_MARKER_END && i < nt) {
                next();
                if (sc_broken) { S.nres = 0; S.nargs = 0; S.result(doc_id, i, 0, 0, 0xFFFFu, 0, 0, OBM_R_HOST, 0); return; }
                S.result(doc_id, marker_tuple, sc_off, sc_len, (uint32_t)def, n_args, arg_base, OBM_R_OK | OBM_R_NL, 0); /* emit.go:8-24 */
                scope_clear(); def = -1; n_args = 0;
                st = S_PARSE;
            } else st = S_PARSE;
            break;
        default: st = S_STOP;
        }
        /* arguments collected for a marker that is abandoned (unknown argument, unexpected lexeme) are dropped */
        if (st == S_PARSE && def >= 0) { S.nargs -= n_args; n_args = 0; def = -1; }
    }
}

} /* namespace obmr */
#endif
