/*
 * obm_corpus.h -- deterministic synthetic manifest generator (BASELINE.md / SURVEY.md section 8d,
 * configs C2/C3/C4).  One document = `doc_bytes` ASCII bytes, generated from
 * splitmix64(0x0B200 ^ global_doc_index), so any shard can be regenerated anywhere:
 *   - "---" line, then 1 resource marker line at the document start
 *       # +operator-builder:resource:field=<ident>,value=<lit>,include[=true|false]
 *   - Kubernetes-style YAML body lines (2-8 space indents, key: value, list items) with
 *       5 inline field markers   "  key: v  # +operator-builder:field:name=<ident[.ident]>,type=<t>[,default=<lit>]"
 *       2 head-comment markers   "  # +operator-builder:field:name=<ident>,type=string,description="<words>""
 *   - "# "-comment filler up to exactly doc_bytes, last byte '\n'
 * flavour 1 writes the collection spelling (+operator-builder:collection:field, collectionField=;
 * reference: internal/workload/v1/markers/collection_field_marker.go:13, manifests/manifest.go:89-95).
 * Documents shorter than 1.5 KiB carry fewer markers (as many as fit).
 *
 * Same source for the device kernel and the host helper: the bytes are identical by construction.
 */
#ifndef OBM_CORPUS_H
#define OBM_CORPUS_H
#include <stdint.h>

#if defined(__CUDACC__)
#define OBMC_HD __host__ __device__
#else
#define OBMC_HD
#endif

namespace obmc {

struct Rng {
    uint64_t s;
    OBMC_HD uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    OBMC_HD uint32_t below(uint32_t n) { return (uint32_t)(next() % n); }
};

struct Writer {
    uint8_t *o; uint32_t n, cap;
    OBMC_HD void ch(char c) { if (n < cap) o[n] = (uint8_t)c; n++; }
    OBMC_HD void str(const char *z) { while (*z) ch(*z++); }
    OBMC_HD void spaces(uint32_t k) { while (k--) ch(' '); }
    OBMC_HD void num(uint32_t v) { char t[12]; int k = 0; do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v); while (k) ch(t[--k]); }
};

OBMC_HD inline const char *word(uint32_t i) {
    switch (i % 24) {
    case 0: return "replicas"; case 1: return "image"; case 2: return "name"; case 3: return "namespace";
    case 4: return "port"; case 5: return "host"; case 6: return "tier"; case 7: return "provider";
    case 8: return "region"; case 9: return "storage"; case 10: return "class"; case 11: return "limit";
    case 12: return "cpu"; case 13: return "memory"; case 14: return "webstore"; case 15: return "ingress";
    case 16: return "service"; case 17: return "label"; case 18: return "app"; case 19: return "config";
    case 20: return "backend"; case 21: return "frontend"; case 22: return "enabled"; default: return "version";
    }
}

/* one ordinary YAML body line (no marker); returns nothing, appends to w */
OBMC_HD inline void body_line(Writer &w, Rng &r) {
    uint32_t indent = 2 * (1 + r.below(4));
    uint32_t k = r.below(12);
    w.spaces(k == 0 ? 0 : indent);
    switch (k) {
    case 0: w.str("apiVersion: apps/v1"); break;
    case 1: w.str("- name: "); w.str(word(r.below(24))); w.ch('-'); w.str(word(r.below(24))); break;
    case 2: w.str(word(r.below(24))); w.ch(':'); break;
    case 3: w.str("image: registry.acme.io/"); w.str(word(r.below(24))); w.str("/nginx:1."); w.num(r.below(30)); break;
    case 4: w.str("containerPort: "); w.num(1024 + r.below(60000)); break;
    case 5: w.str(word(r.below(24))); w.str(": \""); w.str(word(r.below(24))); w.ch('"'); break;
    case 6: w.str("- "); w.str(word(r.below(24))); break;
    case 7: w.str("path: /"); w.str(word(r.below(24))); w.ch('/'); w.str(word(r.below(24))); break;
    case 8: w.str(word(r.below(24))); w.str(": "); w.num(r.below(100000)); break;
    case 9: w.str("kubernetes.io/"); w.str(word(r.below(24))); w.str(": 'true'"); break;
    case 10: w.str(word(r.below(24))); w.str(": "); w.str(word(r.below(24))); w.str("  # plain comment, no marker"); break;
    default: w.str(word(r.below(24))); w.str(": "); w.str(word(r.below(24))); break;
    }
    w.ch('\n');
}

OBMC_HD inline void ident(Writer &w, Rng &r, bool dotted) {
    w.str(word(r.below(24)));
    if (dotted) { uint32_t parts = r.below(3); while (parts--) { w.ch('.'); w.str(word(r.below(24))); } }
}

OBMC_HD inline void field_prefix(Writer &w, int flavour) {
    w.str(flavour ? "+operator-builder:collection:field:name=" : "+operator-builder:field:name=");
}

/* 5 of these per document */
OBMC_HD inline void inline_marker_line(Writer &w, Rng &r, int flavour) {
    w.spaces(2 * (1 + r.below(4)));
    w.str(word(r.below(24))); w.str(": ");
    uint32_t t = r.below(3);
    uint32_t v = r.below(1000);
    if (t == 0) { w.ch('"'); w.str(word(v)); w.ch('"'); } else if (t == 1) w.num(v); else w.str((v & 1) ? "true" : "false");
    w.str("  # "); field_prefix(w, flavour); ident(w, r, true);
    w.str(",type="); w.str(t == 0 ? "string" : t == 1 ? "int" : "bool");
    if (r.below(3)) {
        w.str(",default=");
        if (t == 0) { w.ch('"'); w.str(word(v)); w.ch('"'); } else if (t == 1) w.num(v); else w.str((v & 1) ? "true" : "false");
    }
    w.ch('\n');
}

/* 2 of these per document (head comment above a YAML line) */
OBMC_HD inline void head_marker_line(Writer &w, Rng &r, int flavour) {
    uint32_t indent = 2 * (1 + r.below(4));
    w.spaces(indent); w.str("# "); field_prefix(w, flavour); ident(w, r, false);
    w.str(",type=string,description=\"");
    uint32_t nw = 3 + r.below(6);
    for (uint32_t i = 0; i < nw; i++) { if (i) w.ch(' '); w.str(word(r.below(24))); }
    w.str("\"\n");
    w.spaces(indent); w.str(word(r.below(24))); w.str(": "); w.str(word(r.below(24))); w.ch('\n');
}

OBMC_HD inline void resource_marker_line(Writer &w, Rng &r, int flavour) {
    w.str("# +operator-builder:resource:");
    w.str(flavour ? "collectionField=" : "field=");
    ident(w, r, false);
    w.str(",value=");
    uint32_t t = r.below(3);
    if (t == 0) { w.ch('"'); w.str(word(r.below(24))); w.ch('"'); } else if (t == 1) w.num(r.below(100)); else w.str("true");
    uint32_t inc = r.below(3);
    w.str(inc == 0 ? ",include" : inc == 1 ? ",include=true" : ",include=false");
    w.ch('\n');
}

/* Writes exactly doc_bytes bytes at `out`. */
OBMC_HD inline void generate_doc(uint8_t *out, uint32_t doc_bytes, uint64_t global_doc_index, int flavour) {
    Rng r{0x0B200ull ^ global_doc_index};
    r.next();
    Writer w{out, 0, doc_bytes};
    const uint32_t MARKER_RESERVE = 150; /* longest marker line pair */
    if (doc_bytes >= 4) w.str("---\n");
    if (doc_bytes >= 200) resource_marker_line(w, r, flavour);
    /* how many of the 7 body markers fit */
    uint32_t want = doc_bytes >= 1536 ? 7 : (doc_bytes > 400 ? (doc_bytes - 400) / 170 : 0);
    if (want > 7) want = 7;
    uint32_t emitted = 0;
    /* spread markers: one every `gap` bytes of body */
    uint32_t body_budget = doc_bytes > w.n + 64 ? doc_bytes - w.n - 64 : 0;
    uint32_t gap = want ? body_budget / (want + 1) : 0xFFFFFFFFu;
    uint32_t next_marker_at = w.n + (want ? gap / 2 : 0);
    while (w.n + MARKER_RESERVE + 64 < doc_bytes) {
        if (emitted < want && w.n >= next_marker_at) {
            /* markers 2 and 5 are head-comment markers, the rest inline */
            if (emitted == 2 || emitted == 5) head_marker_line(w, r, flavour); else inline_marker_line(w, r, flavour);
            emitted++;
            next_marker_at += gap;
        } else {
            body_line(w, r);
        }
    }
    /* any markers that did not fit by position (tiny gaps): place them now if room remains */
    while (emitted < want && w.n + MARKER_RESERVE + 8 < doc_bytes) {
        if (emitted == 2 || emitted == 5) head_marker_line(w, r, flavour); else inline_marker_line(w, r, flavour);
        emitted++;
    }
    /* "# " comment filler to the exact size */
    while (w.n < doc_bytes) {
        uint32_t rem = doc_bytes - w.n;
        if (rem == 1) { w.ch('\n'); break; }
        if (rem == 2) { w.str("#\n"); break; }
        uint32_t len = rem > 80 ? 40 + r.below(38) : rem; /* line length incl. '\n' */
        if (rem - len == 1) len -= 1;                     /* never leave a 1-byte remainder that is not '\n'... it is fine, but keep lines >= 2 */
        w.str("# ");
        for (uint32_t i = 2; i + 1 < len; i++) w.ch("filler comment padding "[(i - 2) % 23]);
        w.ch('\n');
    }
}

} /* namespace obmc */
#endif
