/*
 * obm_parse.cpp -- host-side mirror of the lexer's only consumer, internal/markers/parser, running on
 * the GPU tuple stream (through obm_stream_*): SURVEY.md section 8(f) rank 1, the parser token contract
 * of 8(a) a17.
 *
 * Follows parser/{parser,peek,position,consumed,state,definition,emit,error}.go @ 2827f233 state by state:
 *   startParse/parse        state.go:13-46     parseMarkerStart/Scope/Separator   state.go:48-77
 *   parseArg                state.go:79-93     parseArgValue                      state.go:95-153
 *   parseMoreArgs           state.go:155-169   stripQuotes                        state.go:171-175
 *   scopeBuffer/MarkerText  position.go:18, emit.go:8-24      error results       error.go:8-22
 *   registry lookup         definition.go:13-21 (name = scopeBuffer minus the trailing ':')
 * Not modelled (needs the Go struct types behind marker.Define): Argument.SetValue conversion errors
 * (marker/argument.go:91-127) and InflateObject's missing-argument check (marker/marker.go:65-95).
 *
 * Output records (little endian), one per Result, in order:
 *   ok     [u8 0][u32 nlen][marker name][u32 tlen][MarkerText][u32 nargs] { [u32 alen][arg][u8 kind][u32 vlen][value] }
 *   error  [u8 1][u32 mlen][message][u32 tlen][MarkerText]            (an error ends the document, parser.go:63-73)
 * kind: 0 bool, 1 int, 2 float, 3 string.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/obmarkers.h"
#include "go_unicode_tables.h"
#include "obm_core.h"

struct obm_registry {
    struct Def { std::string name; std::vector<std::string> args; };
    std::vector<Def> defs;
};

extern "C" obm_registry *obm_registry_new(void) { return new (std::nothrow) obm_registry(); }
extern "C" void obm_registry_free(obm_registry *r) { delete r; }
extern "C" int obm_registry_add(obm_registry *r, const char *marker_name, const char *const *arg_names, uint32_t nargs) {
    if (!r || !marker_name) return OBM_E_ARG;
    obm_registry::Def d; d.name = marker_name;
    for (uint32_t i = 0; i < nargs; i++) d.args.push_back(arg_names[i]);
    r->defs.push_back(d);
    return OBM_OK;
}
/* The three markers operator-builder registers: internal/workload/v1/markers/field_marker.go:19,26-38,
 * collection_field_marker.go:13,22, resource_marker.go:25,47-57 (argument names = lowerCamelCase field names). */
extern "C" obm_registry *obm_registry_operator_builder(void) {
    obm_registry *r = obm_registry_new();
    if (!r) return nullptr;
    const char *field[] = {"name", "type", "description", "default", "replace"};
    const char *res[] = {"field", "collectionField", "value", "include"};
    obm_registry_add(r, "+operator-builder:field", field, 5);
    obm_registry_add(r, "+operator-builder:collection:field", field, 5);
    obm_registry_add(r, "+operator-builder:resource", res, 4);
    return r;
}

/* names of the registry (internal use by the device index, obm_lib.cu) */
extern "C" uint32_t obm_registry_names(const obm_registry *r, const char **names, uint32_t *lens, uint32_t cap) {
    if (cap == 0xFFFFFFFFu) return (uint32_t)r->defs.size(); /* count only */
    uint32_t n = 0;
    for (const auto &d : r->defs) { if (n >= cap) break; names[n] = d.name.data(); lens[n] = (uint32_t)d.name.size(); n++; }
    return n;
}

namespace {

struct Lx { int32_t type; std::string value; int64_t line, col; };

void put_u32(std::string &o, uint32_t v) { o.append((const char *)&v, 4); }

/* strconv.Quote of the byte strings that can reach a strconv error here */
std::string go_quote_simple(const std::string &s) {
    static const char hex[] = "0123456789abcdef";
    std::string o = "\"";
    for (unsigned char c : s) {
        if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back((char)c); }
        else if (c == '\n') o += "\\n"; else if (c == '\t') o += "\\t"; else if (c == '\r') o += "\\r";
        else if (c == '\v') o += "\\v"; else if (c == '\f') o += "\\f";
        else if (c < 0x20) { o += "\\x"; o.push_back(hex[c >> 4]); o.push_back(hex[c & 15]); }
        else if (c == 0x7F) o += "\\u007f";
        else o.push_back((char)c);
    }
    o.push_back('"');
    return o;
}

bool parse_bool_ok(const std::string &v) { /* strconv.ParseBool */
    static const char *ok[] = {"1", "t", "T", "TRUE", "true", "True", "0", "f", "F", "FALSE", "false", "False"};
    for (const char *k : ok) if (v == k) return true;
    return false;
}

/* does the (lexer-validated) decimal literal overflow float32?  value >= 2^128 - 2^103 = 3.4028235677973366e38 */
bool float32_overflows(const std::string &v) {
    static const char *T = "340282356779733661637539395458142568448"; /* 2^128 - 2^103, 39 digits */
    size_t i = 0;
    if (i < v.size() && (v[i] == '+' || v[i] == '-')) i++;
    std::string digits; long long dp = 0; bool dot = false, any = false;
    for (; i < v.size(); i++) {
        char c = v[i];
        if (c == '.') { dot = true; continue; }
        if (c < '0' || c > '9') break;
        if (digits.empty() && c == '0') { if (dot) dp--; continue; }
        digits.push_back(c); any = true;
        if (!dot) dp++;
    }
    if (!any) return false;
    if (i < v.size() && (v[i] == 'e' || v[i] == 'E')) {
        i++; int sg = 1; long long e = 0;
        if (i < v.size() && (v[i] == '+' || v[i] == '-')) { if (v[i] == '-') sg = -1; i++; }
        for (; i < v.size() && v[i] >= '0' && v[i] <= '9'; i++) if (e < 10000) e = e * 10 + (v[i] - '0');
        dp += sg * e;
    }
    /* value = 0.d1d2... * 10^dp */
    if (dp > 39) return true;
    if (dp < 39) return false;
    for (size_t k = 0; k < 39; k++) {
        char c = k < digits.size() ? digits[k] : '0';
        if (c > T[k]) return true;
        if (c < T[k]) return false;
    }
    return true;
}

struct Parser {
    obm_stream *s; const obm_registry *reg;
    std::string scope;             /* scopeBuffer */
    Lx stack[3]; int peek_count = 0;
    Lx current;
    const obm_registry::Def *def = nullptr;
    struct Arg { std::string name; uint8_t kind; std::string value; };
    std::vector<Arg> args;
    std::string out; int64_t nresults = 0;

    Lx next_lexeme() {
        obm_lexeme l; Lx r;
        if (obm_stream_next(s, &l)) { r.type = l.type; r.value.assign((const char *)l.value, l.value_len); r.line = l.line; r.col = l.column; }
        else { r.type = 0; r.line = r.col = 0; }
        return r;
    }
    const Lx &peek() { /* peek.go:8-22 */
        if (peek_count > 0) return stack[peek_count - 1];
        peek_count = 1; stack[2] = stack[1]; stack[1] = stack[0]; stack[0] = next_lexeme();
        return stack[0];
    }
    bool peeked(int t) { return peek().type == t; }
    void next() { /* position.go:7-20 */
        if (peek_count > 0) peek_count--; else { stack[2] = stack[1]; stack[1] = stack[0]; stack[0] = next_lexeme(); }
        scope += stack[peek_count].value; current = stack[peek_count];
    }
    void discard() { /* position.go:23-31 */
        if (peek_count > 1) for (int i = peek_count < 3 ? peek_count : 2; i > 0; i--) stack[i] = stack[i - 1];
        stack[0] = next_lexeme();
    }
    bool consumed(int t) { if (peek().type == t) { next(); return true; } return false; }
    void flush() { scope.clear(); def = nullptr; }
    void error(const std::string &msg) { /* error.go:8-22 */
        char pos[96]; snprintf(pos, sizeof pos, "{line:%lld column:%lld}", (long long)current.line, (long long)current.col);
        std::string m = msg + ", on marker " + (def ? def->name : std::string("Unknown Marker")) + " at " + pos;
        out.push_back(1); put_u32(out, (uint32_t)m.size()); out += m; put_u32(out, (uint32_t)scope.size()); out += scope;
        nresults++;
    }
    void emit() { /* emit.go:8-24 */
        out.push_back(0); put_u32(out, (uint32_t)def->name.size()); out += def->name; put_u32(out, (uint32_t)scope.size()); out += scope;
        put_u32(out, (uint32_t)args.size());
        for (const Arg &a : args) { put_u32(out, (uint32_t)a.name.size()); out += a.name; out.push_back((char)a.kind); put_u32(out, (uint32_t)a.value.size()); out += a.value; }
        nresults++;
        flush();
    }
    bool lookup_arg(const std::string &n) const { for (const auto &a : def->args) if (a == n) return true; return false; }

    enum St { S_START, S_PARSE, S_MARKER_START, S_SCOPE, S_SEPARATOR, S_ARG, S_MORE, S_STOP };
    void run() {
        St st = S_START;
        while (st != S_STOP) {
            switch (st) {
            case S_START: case S_PARSE: /* state.go:13-46 */
                if (peeked(OBM_K_COMMENT)) { discard(); st = S_PARSE; }
                else if (consumed(OBM_K_MARKER_START)) st = S_MARKER_START;
                else if (consumed(OBM_K_EOF)) st = S_STOP;
                else if (st == S_PARSE && consumed(OBM_K_ERROR)) { error(current.value); st = S_STOP; }
                else if (st == S_START) st = S_PARSE;
                else { next(); scope.clear(); st = S_PARSE; }
                break;
            case S_MARKER_START: st = consumed(OBM_K_SCOPE) ? S_SCOPE : S_PARSE; break;
            case S_SCOPE: st = consumed(OBM_K_SEPARATOR) ? S_SEPARATOR : S_PARSE; break;
            case S_SEPARATOR: /* state.go:64-77 */
                if (consumed(OBM_K_SCOPE)) { st = S_SCOPE; break; }
                if (peeked(OBM_K_ARG) && !scope.empty()) {
                    std::string name = scope.substr(0, scope.size() - 1);
                    const obm_registry::Def *d = nullptr;
                    for (const auto &x : reg->defs) if (x.name == name) { d = &x; break; }
                    if (d) { def = d; args.clear(); st = S_ARG; break; }
                }
                flush(); st = S_PARSE;
                break;
            case S_ARG: { /* state.go:79-93 + parseArgValue :95-153 */
                if (!consumed(OBM_K_ARG) || !lookup_arg(current.value)) { st = S_PARSE; break; }
                std::string arg = current.value;
                if (peeked(OBM_K_ARG_ASSIGNMENT)) next();
                if (peeked(OBM_K_QUOTE)) next();
                if (peeked(OBM_K_SYNTHETIC_BOOL)) {
                    std::string v = peek().value;
                    if (!parse_bool_ok(v)) { error("strconv.ParseBool: parsing " + go_quote_simple(v) + ": invalid syntax"); st = S_STOP; break; }
                    args.push_back(Arg{arg, 0, v}); discard();
                } else if (consumed(OBM_K_BOOL_LITERAL)) {
                    if (!parse_bool_ok(current.value)) { error("strconv.ParseBool: parsing " + go_quote_simple(current.value) + ": invalid syntax"); st = S_STOP; break; }
                    args.push_back(Arg{arg, 0, current.value});
                } else if (consumed(OBM_K_INTEGER_LITERAL)) {
                    args.push_back(Arg{arg, 1, current.value});
                } else if (consumed(OBM_K_FLOAT_LITERAL)) {
                    if (float32_overflows(current.value)) { error("strconv.ParseFloat: parsing " + go_quote_simple(current.value) + ": value out of range"); st = S_STOP; break; }
                    args.push_back(Arg{arg, 2, current.value});
                } else if (consumed(OBM_K_STRING_LITERAL)) {
                    args.push_back(Arg{arg, 3, current.value});
                    if (peeked(OBM_K_QUOTE)) next();
                } else { st = S_PARSE; break; }
                st = S_MORE;
                break;
            }
            case S_MORE: /* state.go:155-169 */
                if (consumed(OBM_K_ARG_DELIMITER)) st = S_ARG;
                else if (consumed(OBM_K_MARKER_END)) { emit(); st = S_PARSE; }
                else st = S_PARSE;
                break;
            default: st = S_STOP;
            }
        }
    }
};

} // namespace

/* registry -> the flat form the device walk uses (csrc/obm_parse_dev.h); false when it does not fit */
#include "obm_parse_dev.h"
extern "C" bool obm_registry_flatten(const obm_registry *r, obmr::DevRegistry *D) {
    memset(D, 0, sizeof *D);
    if (r->defs.size() > 8) return false;
    uint32_t o = 0, a = 0;
    for (size_t i = 0; i < r->defs.size(); i++) {
        const auto &d = r->defs[i];
        if (o + d.name.size() > sizeof D->text) return false;
        D->name_off[i] = o; memcpy(D->text + o, d.name.data(), d.name.size()); o += (uint32_t)d.name.size();
    }
    D->name_off[r->defs.size()] = o; D->n = (uint32_t)r->defs.size();
    for (size_t i = 0; i < r->defs.size(); i++) {
        D->arg_first[i] = a;
        for (const auto &n : r->defs[i].args) {
            if (a >= 64 || o + n.size() > sizeof D->text) return false;
            D->arg_off[a++] = o; memcpy(D->text + o, n.data(), n.size()); o += (uint32_t)n.size();
        }
    }
    D->arg_first[r->defs.size()] = a; D->arg_off[a] = o;
    return true;
}

/* host run of the device walk over one document (test / mirror use: the records the GPU would write) */
extern "C" int64_t obm_parse_doc_records(const obm_registry *reg, const uint8_t *doc, const obm_tuple *tuples, uint64_t ntuples, uint32_t doc_id,
                                         obm_result *res, uint64_t res_cap, obm_arg *args, uint64_t arg_cap, uint64_t *nargs_out) {
    obmr::DevRegistry D;
    if (!reg || !obm_registry_flatten(reg, &D)) return OBM_E_ARG;
    obmr::Sink S{res, res_cap, args, arg_cap, 0, 0, 0, 0};
    obmr::parse_doc(D, doc, tuples, (uint32_t)ntuples, doc_id, S);
    if (nargs_out) *nargs_out = S.nargs;
    return S.nres;
}

extern "C" int64_t obm_results_format_doc(const obm_registry *reg, const uint8_t *doc, uint64_t doc_len, const obm_tuple *tuples, uint64_t ntuples,
                                          const obm_result *results, uint64_t nresults, const obm_arg *args_base, uint8_t **out, uint64_t *out_len) {
    if (!reg || !out || !out_len) return OBM_E_ARG;
    if (nresults == 1 && (results[0].flags & OBM_R_HOST)) return obm_parse_doc(reg, doc, doc_len, tuples, ntuples, out, out_len);
    std::string o;
    for (uint64_t i = 0; i < nresults; i++) {
        const obm_result &r = results[i];
        if (r.reg_id >= reg->defs.size()) return OBM_E_ARG;
        const obm_registry::Def &d = reg->defs[r.reg_id];
        std::string text((const char *)doc + r.text_off, r.text_len);
        if (r.flags & OBM_R_NL) text.push_back('\n');
        if (r.flags & OBM_R_OK) {
            o.push_back(0); put_u32(o, (uint32_t)d.name.size()); o += d.name; put_u32(o, (uint32_t)text.size()); o += text;
            put_u32(o, r.nargs);
            for (uint32_t a = 0; a < r.nargs; a++) {
                const obm_arg &g = args_base[r.arg_base + a];
                put_u32(o, g.name_len); o.append((const char *)doc + g.name_off, g.name_len); o.push_back((char)g.kind);
                if (g.flags & OBM_A_SYNTHETIC_TRUE) { put_u32(o, 4); o += "true"; }
                else { put_u32(o, g.val_len); o.append((const char *)doc + g.val_off, g.val_len); }
            }
        } else { /* error.go:8-22: "<msg>, on marker <name> at {line:L column:C}" */
            const obm_arg &g = args_base[r.arg_base];
            const std::string v((const char *)doc + g.val_off, g.val_len);
            std::string m = (r.flags & OBM_R_ERR_PARSEBOOL) ? "strconv.ParseBool: parsing " + go_quote_simple(v) + ": invalid syntax"
                                                             : "strconv.ParseFloat: parsing " + go_quote_simple(v) + ": value out of range";
            char pos[96]; snprintf(pos, sizeof pos, "{line:%u column:%u}", g.name_off, r.aux);
            m += ", on marker " + d.name + " at " + pos;
            o.push_back(1); put_u32(o, (uint32_t)m.size()); o += m; put_u32(o, (uint32_t)text.size()); o += text;
        }
    }
    uint8_t *buf = (uint8_t *)malloc(o.size() ? o.size() : 1);
    if (!buf) return OBM_E_NOMEM;
    memcpy(buf, o.data(), o.size());
    *out = buf; *out_len = o.size();
    return (int64_t)nresults;
}

extern "C" int64_t obm_parse_doc(const obm_registry *reg, const uint8_t *doc, uint64_t doc_len, const obm_tuple *tuples, uint64_t ntuples,
                                 uint8_t **out, uint64_t *out_len) {
    if (!reg || !out || !out_len) return OBM_E_ARG;
    obm_stream *s = obm_stream_new(doc, doc_len, tuples, ntuples);
    if (!s) return OBM_E_NOMEM;
    Parser p; p.s = s; p.reg = reg;
    p.run();
    obm_stream_free(s);
    uint8_t *buf = (uint8_t *)malloc(p.out.size() ? p.out.size() : 1);
    if (!buf) return OBM_E_NOMEM;
    memcpy(buf, p.out.data(), p.out.size());
    *out = buf; *out_len = p.out.size();
    return p.nresults;
}
