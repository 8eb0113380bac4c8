/*
 * obm_pipe.cuh -- kernels of the two-stage pipeline (mode 0); logic in obm_pipe.h / obm_tile.h.
 *
 *   k_tile_units  units (K1 sub-batches: <= DMAX whole documents of one 16 KiB tile) per tile; an exclusive
 *                 scan gives every unit a static id in document order
 *   k1_scan       tile-resident classification + bit-parallel line logic (no lexing).  Per unit: one item per
 *                 tuple-owning line in position order, every document closed by an EOF item, plus a 16-byte
 *                 unit record.  No cross-CTA dependency.
 *   k2_units      ONE WARP per unit, no block barriers: marker items are compacted (ballots) and lexed ONCE,
 *                 one lane per line, tuples staged in the warp's shared-memory area; item counts -> unit total
 *                 -> decoupled look-back over units -> final positions; comment / EOF tuples are written in
 *                 place, staged marker tuples copied out lane-per-tuple.  Lines that do not fit the staging area
 *                 are lexed again straight to their final place; documents whose lines interact (or that K1
 *                 flagged: non-ASCII, too many owning lines) are lexed by the exact Unicode instantiation, all
 *                 inside the warp.  Writes doc_tuple_off as it goes.
 */
#pragma once
#include "obm_fast.cuh"
#include "obm_pipe.h"

namespace obmq {

using obmt::SmemScan;
using obmp::item_t;

struct PipeArgs {
    const uint8_t *bytes; const uint64_t *doc_off; uint32_t ndocs; uint64_t total_bytes;
    const uint32_t *tile_first; uint32_t ntiles;
    const uint64_t *ubase;      /* [ntiles+1] exclusive scan of units per tile; [ntiles] = number of units */
    const obmp::TileRec *trec;  /* [ntiles] */
    /* K1 -> K2 */
    item_t *items; uint64_t items_cap;
    obmp::Unit *units;          /* [nunits] */
    uint32_t *doc_flag;
    uint64_t *st_tuples, *st_blocks; /* two-level look-back chain over units (obmf::lookback2_warp) */
    /* results */
    uint32_t *counts; obm_tuple *out; uint64_t out_cap; uint64_t *tuple_off;
    uint32_t *status; unsigned long long *totals;
    uint32_t *ctl;
};
enum { CT_T1 = 0, CT_T2 = 1, CT_UNI = 2 /* some document needs per-line Unicode lexing */, CT_ITOP = 4 /* u64 at ctl[4..5] */, CT_OVF = 6 };

__device__ __forceinline__ obm::Tables dev_tables() {
    obm::Tables T;
    T.letter = D_GO_LETTER_RANGES; T.n_letter = D_GO_LETTER_RANGES_N;
    T.number = D_GO_NUMBER_RANGES; T.n_number = D_GO_NUMBER_RANGES_N;
    T.f64_overflow_digits = D_F64_OVERFLOW_DIGITS;
    return T;
}

__global__ void __launch_bounds__(256)
k_tile_units(const uint64_t *__restrict__ doc_off, const uint32_t *__restrict__ tile_first, uint32_t ntiles, uint32_t *__restrict__ nsub,
             obmp::TileRec *__restrict__ trec) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint32_t d0 = tile_first[t], d1 = tile_first[t + 1];
    uint32_t n = 0;
    obmp::TileRec r{d0, d1, 0, 0, 0, 0};
    if (d1 > d0) {
        const uint32_t large = (doc_off[d1] - doc_off[d1 - 1] > obmt::MAXDOC) ? 1u : 0u;
        const uint32_t ns = d1 - d0 - large;
        n = (ns + obmt::DMAX - 1) / obmt::DMAX;
        if (n == 0) n = 1;
        r.d_last = d1 | (large << 31);
        r.b0 = doc_off[d0]; r.b1 = doc_off[d0 + (ns < obmt::DMAX ? ns : obmt::DMAX)];
    }
    nsub[t] = n;
    trec[t] = r;
}

/* ---------------------------------------------------------------------------------------------- K1 -- */
struct K1Shared {
    SmemScan S;
    alignas(8) item_t sitems[obmt::QMAX]; /* items of the sub-batch in owner order */
    uint16_t dlast[obmt::DMAX + 1];       /* index after the last owner of document k */
    alignas(8) uint64_t mbar;
    uint64_t item_base;
    obmp::TileRec rec;                    /* the NEXT tile's record (prefetched while this one is processed) */
    uint64_t u0;
    uint32_t tile, pre_issued;
};

__global__ void __launch_bounds__(obmt::NT, 4)
k1_scan(PipeArgs A) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    K1Shared &C = *reinterpret_cast<K1Shared *>(smem_raw);
    SmemScan &S = C.S;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { obmf::mbar_init(&C.mbar, 1); obmf::fence_mbar_init(); }
    __syncthreads();
    uint32_t mbar_phase = 0;
    item_t *sitems = C.sitems;
    /* software pipeline over tiles: the ticket of the next tile is taken when this one starts, its record is
     * fetched in the middle, and its first TMA load is issued as soon as this tile is done with S.data */
    if (tid == 0) C.tile = atomicAdd(&A.ctl[CT_T1], 1u);
    __syncthreads();
    uint32_t t = C.tile;
    obmp::TileRec rec{0, 0, 0, 0, 0, 0}; uint64_t u0 = 0; bool pre_issued = false;
    if (t < A.ntiles) { rec = A.trec[t]; u0 = A.ubase[t]; }
    __syncthreads();
    for (;;) {
        if (t >= A.ntiles) break;
        uint32_t nt = 0; obmp::TileRec nrec{0, 0, 0, 0, 0, 0}; uint64_t nu0 = 0; bool fetched = false, issued_next = false;
        if (tid == 0) nt = atomicAdd(&A.ctl[CT_T1], 1u);
        auto fetch_next = [&]() { if (!fetched) { fetched = true; if (nt < A.ntiles) { nrec = A.trec[nt]; nu0 = A.ubase[nt]; } } };
        auto publish_next = [&]() { /* thread 0, before the barrier that ends the tile */
            fetch_next();
            C.tile = nt; C.rec = nrec; C.u0 = nu0; C.pre_issued = issued_next ? 1u : 0u;
        };
        const uint32_t d_first = rec.d_first, d_last = rec.d_last & 0x7FFFFFFFu;
        const bool has_large = (rec.d_last >> 31) != 0;
        if (d_last == d_first) {
            if (tid == 0) publish_next();
            __syncthreads();
            t = C.tile; rec = C.rec; u0 = C.u0; pre_issued = false;
            __syncthreads();
            continue;
        }
        const uint32_t d_small_end = d_last - (has_large ? 1u : 0u);
        uint32_t nsub = (d_small_end - d_first + obmt::DMAX - 1) / obmt::DMAX;
        if (nsub == 0) nsub = 1;
        if (has_large && tid == 0) A.doc_flag[d_last - 1] = obmp::GF_LARGE;
        for (uint32_t k = 0; k < nsub; k++) {
            const uint32_t da = d_first + k * obmt::DMAX, db = min(da + obmt::DMAX, d_small_end), nd = db - da;
            const uint32_t extra = (k == nsub - 1 && has_large) ? 1u : 0u;
            const uint64_t u = u0 + k;
            uint32_t n_owners = 0, n_live = 0;
            if (nd) {
                const uint64_t b0 = k == 0 ? rec.b0 : A.doc_off[da], b1 = k == 0 ? rec.b1 : A.doc_off[db];
                const uint64_t abs0 = (uint64_t)(uintptr_t)A.bytes + b0, base_abs = abs0 & ~15ull;
                const uint32_t skew = (uint32_t)(abs0 - base_abs), span = (uint32_t)(b1 - b0) + skew, load = (span + 15u) & ~15u;
                if (tid == 0) {
                    S.nd = nd; S.lo_pos = skew; S.hi_pos = span; S.n_owners = 0;
                    if (load && !(k == 0 && pre_issued)) { obmf::fence_proxy_async(); obmf::mbar_expect_tx(&C.mbar, load); obmf::tma_bulk_g2s(S.data, (const void *)(uintptr_t)base_abs, load, &C.mbar); }
                }
                if (tid <= nd) S.dstart[tid] = (uint32_t)(A.doc_off[da + tid] - b0) + skew;
                __syncthreads();
                if (load) { obmf::mbar_wait(&C.mbar, mbar_phase); mbar_phase ^= 1; }
                /* P2 classify */
                const uint32_t nwords = (span + 31) >> 5, nwr = (nwords + 31u) & ~31u;
                for (uint32_t base = 0; base < nwords; base += obmt::NT) { uint32_t wi = base + tid; if (wi < nwr && wi < obmt::NW) obmt::classify_word(S, wi); }
                for (uint32_t wi = nwr + tid; wi < obmt::NW; wi += obmt::NT) { S.nlw[wi] = 0; S.spw[wi] = 0; }
                __syncthreads();
                /* P3 doc prep */
                if (tid < nd) obmt::doc_prep(S, tid, true);
                if (tid == 0) fetch_next(); /* the ticket taken at the top has long arrived; the record is used at the end of the tile */
                __syncthreads();
                /* P4 bit-parallel line scan */
                uint32_t nl[obmt::WPT], sp[obmt::WPT], lm[obmt::WPT];
                {
                    const uint4 a = reinterpret_cast<const uint4 *>(S.nlw)[tid], b = reinterpret_cast<const uint4 *>(S.spw)[tid];
                    nl[0] = a.x; nl[1] = a.y; nl[2] = a.z; nl[3] = a.w; sp[0] = b.x; sp[1] = b.y; sp[2] = b.z; sp[3] = b.w;
                }
                obmt::line_starts(S, tid, nl, lm);
                obmt::LineBits lb;
                {
                    const uint32_t lane = tid & 31, wid = tid >> 5;
                    const uint32_t c0 = obmt::first_events(nl, sp, lm, 0, nullptr), c1 = obmt::first_events(nl, sp, lm, 1, nullptr);
                    const uint32_t Gb = __ballot_sync(0xffffffffu, c0 != 0), Pb = __ballot_sync(0xffffffffu, c1 != 0 && c0 == 0);
                    uint32_t w0, w1;
                    obmt::carry_lookahead32(Gb, Pb, 0, &w0);
                    obmt::carry_lookahead32(Gb, Pb, 1, &w1);
                    if (lane == 0) S.scan_tmp[wid] = w0 | ((w1 & ~w0 & 1u) << 1);
                    __syncthreads();
                    uint32_t cin = 0;
                    for (uint32_t w = 0; w < wid; w++) { uint32_t f = S.scan_tmp[w]; cin = (f & 1u) | ((f >> 1) & cin); }
                    __syncthreads();
                    uint32_t dummy;
                    const uint32_t Cm = obmt::carry_lookahead32(Gb, Pb, cin, &dummy);
                    obmt::first_events(nl, sp, lm, (Cm >> lane) & 1u, &lb);
                }
                uint32_t my_owners = 0, my_nl = 0;
#pragma unroll
                for (uint32_t j = 0; j < obmt::WPT; j++) { my_owners += (uint32_t)__popc(lb.own[j]); my_nl += (uint32_t)__popc(nl[j]); }
                uint32_t tot;
                uint32_t pre = obmf::block_scan_excl(my_nl | (my_owners << 16), S.scan_tmp, tot);
                n_owners = tot >> 16;
                {
                    uint32_t nlp = pre & 0xFFFFu, own = pre >> 16;
#pragma unroll
                    for (uint32_t j = 0; j < obmt::WPT; j++) { S.nlpre[tid * obmt::WPT + j] = (uint16_t)nlp; nlp += (uint32_t)__popc(nl[j]); }
                    if (n_owners <= obmt::QMAX) {
#pragma unroll
                        for (uint32_t j = 0; j < obmt::WPT; j++) {
                            uint32_t bits = lb.own[j];
                            while (bits) { S.owner[own++] = (tid * obmt::WPT + j) * 32 + (uint32_t)(__ffs((int)bits) - 1); bits &= bits - 1; }
                        }
                    } else {
                        n_owners = 0;
                        if (tid < nd) S.dflag[tid] |= obmt::DF_QOVERFLOW;
                    }
                }
                __syncthreads();
                /* P5 owners -> items (shared memory, owner order); lines that own no tuple ("dead": a quote or slash
                 * outside any comment) are dropped here: S.owner[o] becomes the number of live owners before o */
                uint32_t live_run = 0;
                for (uint32_t o0 = 0; o0 < n_owners; o0 += obmt::NT) {
                    const uint32_t o = o0 + tid;
                    bool live = false;
                    if (o < n_owners) { const item_t it = obmp::k1_owner_item(S, o); sitems[o] = it; live = !obmp::it_dead(it); }
                    const uint32_t bal = __ballot_sync(0xffffffffu, live);
                    if ((tid & 31) == 0) S.scan_tmp[tid >> 5] = (uint32_t)__popc(bal);
                    __syncthreads();
                    uint32_t pre = 0, tot = 0;
#pragma unroll
                    for (uint32_t w = 0; w < obmt::NT / 32; w++) { const uint32_t c = S.scan_tmp[w]; if (w < (tid >> 5)) pre += c; tot += c; }
                    if (o < n_owners) S.owner[o] = live_run + pre + (uint32_t)__popc(bal & ((1u << (tid & 31)) - 1u));
                    live_run += tot;
                    __syncthreads();
                }
                n_live = live_run;
                /* S.data is dead from here on: issue the next tile's first load now, under the rest of this tile */
                if (tid == 0 && k == nsub - 1) {
                    fetch_next();
                    if (nt < A.ntiles && (nrec.d_last & 0x7FFFFFFFu) > nrec.d_first && nrec.b1 > nrec.b0) {
                        const uint64_t nabs0 = (uint64_t)(uintptr_t)A.bytes + nrec.b0, nbase = nabs0 & ~15ull;
                        const uint32_t nload = ((uint32_t)(nrec.b1 - nrec.b0) + (uint32_t)(nabs0 - nbase) + 15u) & ~15u;
                        obmf::fence_proxy_async(); obmf::mbar_expect_tx(&C.mbar, nload); obmf::tma_bulk_g2s(S.data, (const void *)(uintptr_t)nbase, nload, &C.mbar);
                        issued_next = true;
                    }
                }
                /* per document: index after its last owner (owners are in position order, hence grouped by document) */
                if (tid < nd) {
                    uint32_t lo = 0, hi = n_owners;
                    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (obmp::it_doc(sitems[mid]) <= tid) lo = mid + 1; else hi = mid; }
                    C.dlast[tid] = (uint16_t)(lo < n_owners ? S.owner[lo] : n_live); /* live items before the next document */
                    const uint32_t f = S.dflag[tid];
                    A.doc_flag[da + tid] = ((f & obmt::DF_NONASCII) ? obmp::GF_NONASCII : 0u) | ((f & obmt::DF_QOVERFLOW) ? obmp::GF_QOVERFLOW : 0u);
                    if (f & obmt::DF_UNI) A.ctl[CT_UNI] = 1;
                }
            }
            __syncthreads();
            /* the unit's items are contiguous; units are placed by a bump allocator (any order) */
            const uint32_t n_items = n_live + nd + extra;
            if (tid == 0) C.item_base = atomicAdd(reinterpret_cast<unsigned long long *>(&A.ctl[CT_ITOP]), (unsigned long long)n_items);
            __syncthreads();
            const uint64_t ibase = C.item_base;
            const bool room = ibase + n_items <= A.items_cap;
            if (room) {
                for (uint32_t o = tid; o < n_owners; o += obmt::NT) { const item_t it = sitems[o]; if (!obmp::it_dead(it)) A.items[ibase + S.owner[o] + obmp::it_doc(it)] = it; }
                if (tid < nd) A.items[ibase + C.dlast[tid] + tid] = obmp::make_eof_item(S.dstart[tid + 1] - S.dstart[tid], tid, (S.dflag[tid] & obmt::DF_EXACT_MASK) != 0);
                if (extra && tid == 0) A.items[ibase + n_items - 1] = obmp::make_large_item();
            }
            if (tid == 0) {
                if (!room) A.ctl[CT_OVF] = 1;
                A.units[u] = obmp::Unit{ibase, da, n_items | (nd << 16)};
                if (k == nsub - 1) publish_next();
            }
            __syncthreads();
        }
        t = C.tile; rec = C.rec; u0 = C.u0; pre_issued = C.pre_issued != 0;
        __syncthreads();
    }
}

/* ---------------------------------------------------------------------------------------------- K2 -- */
struct K2Warp {
    alignas(16) obm_tuple stage[obmp::W_MLCAP * obmp::W_LTS];
    uint64_t moff[obmp::W_MLCAP];   /* staged line -> final output position (~0: not copied) */
    alignas(16) uint8_t pool[obmp::W_POOL * 16]; /* staged line text (obm_pipe.h: LineView) */
    uint16_t icnt[obmp::W_ICAP];    /* tuples per item of the block (G_CNT_LOOKUP: counts[doc]) */
    uint8_t mlist[obmp::W_ICAP];    /* marker rank -> item index inside the block */
};

struct K2Ctx {
    const PipeArgs &A; K2Warp &C; const obm::Tables &T;
    uint64_t i0; uint32_t d0, nd, lane; bool writing;
    uint32_t markers, lexemes, exact, fatal;
    __device__ __forceinline__ uint32_t doc_of(item_t it) const { return obmp::it_large(it) ? d0 + nd : d0 + obmp::it_doc(it); }
};

/* lanes with uni == true: a line with bytes >= 0x80 of a valid-UTF-8 document, lexed from its start by the Unicode
 * lexer straight from global memory.  Kept apart from k2_lex_lines so that this rare call does not weigh on the
 * register allocation of the hot path.  Whole warp must call. */
__device__ __forceinline__ uint32_t k2_lex_uni_lines(K2Ctx &X, bool uni, item_t it, uint32_t d, obm_tuple *out, uint32_t cap, uint32_t *mk, uint32_t *lx) {
    uint32_t r = 0;
    if (uni) {
        const uint64_t o0 = X.A.doc_off[d];
        r = obmp::k2_unicode_item(X.T, X.A.bytes + o0, (uint32_t)(X.A.doc_off[d + 1] - o0), it, out, cap, mk, lx);
    }
    __syncwarp();
    return r;
}

/* Lanes with on == true lex the marker line of item `it` (document d) into out[0..cap) (cap 0: count only).
 * The lines' text is packed into the warp's pool (exclusive scan of 16-byte chunk counts, cp.async for all of
 * them, one wait) and lexed through the shared window; a line that does not fit, or whose lookahead could
 * run past its copy (obm_pipe.h: line_view_safe), is lexed from global memory.  Whole warp must call. */
__device__ __forceinline__ uint32_t k2_lex_lines(K2Ctx &X, bool on, item_t it, uint32_t d, obm_tuple *out, uint32_t cap, uint32_t *mk, uint32_t *lx) {
    K2Warp &C = X.C; const PipeArgs &A = X.A; const uint32_t lane = X.lane;
    uint32_t len = 0; const uint8_t *gdoc = nullptr;
    obmp::LineView v{0, 0, 0};
    if (on) {
        const uint64_t o0 = A.doc_off[d];
        len = (uint32_t)(A.doc_off[d + 1] - o0); gdoc = A.bytes + o0;
        v = obmp::line_view(gdoc, len, it, A.bytes, A.total_bytes);
    }
    const uint32_t want = (on && v.nch <= 32u) ? v.nch : 0u;
    uint32_t incl = want;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
    const bool fits = want != 0 && incl <= obmp::W_POOL;
    const uint32_t off = incl - want;
    uint32_t todo = __ballot_sync(0xffffffffu, fits);
    while (todo) {
        const uint32_t q = (uint32_t)__ffs((int)todo) - 1u; todo &= todo - 1u;
        const uint32_t nq = __shfl_sync(0xffffffffu, want, q), oq = __shfl_sync(0xffffffffu, off, q);
        const unsigned long long gq = __shfl_sync(0xffffffffu, (unsigned long long)v.g0, q);
        if (lane < nq) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(C.pool + (size_t)(oq + lane) * 16u);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gq + (unsigned long long)lane * 16ull) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    uint32_t r = 0;
    if (on) {
        uint32_t nv = 0; const uint8_t *sm = C.pool + (size_t)off * 16u;
        if (fits) nv = obmp::line_view_safe(sm, v, gdoc, len, it);
        if (nv) {
            const uint8_t *doc = sm + (intptr_t)((uintptr_t)gdoc - v.g0);
            r = obmp::k2_marker_item(X.T, obm::ShBytes{(uint32_t)__cvta_generic_to_shared(C.pool) + (uint32_t)(doc - C.pool), doc}, nv, it, out, cap, mk, lx);
        } else r = obmp::k2_marker_item(X.T, gdoc, len, it, out, cap, mk, lx);
    }
    __syncwarp(); /* the pool is reused by the next round */
    return r;
}

/* Block [b0, b1) of the unit's items: compact the marker items, lex each line once (lane per line, the first
 * W_MLCAP lines stage their tuples), counts of everything else.  stable: document flags are final (large
 * units, second sweep) -- lines of flagged documents are skipped.  Returns lane-local "needs the flag pass". */
template <bool UNI>
__device__ __forceinline__ bool k2_lex_block(K2Ctx &X, uint32_t b0, uint32_t b1, bool stable, uint32_t &n_ml_out) {
    K2Warp &C = X.C; const PipeArgs &A = X.A; const uint32_t lane = X.lane;
    uint32_t n_ml = 0; bool any = false;
    C.moff[lane] = ~0ull;
    for (uint32_t c0 = b0; c0 < b1; c0 += 32) {
        const uint32_t i = c0 + lane; const bool valid = i < b1;
        const item_t it = valid ? A.items[X.i0 + i] : 0;
        bool m = valid && obmp::it_marker(it);
        if (stable && valid && A.doc_flag[X.doc_of(it)]) { m = false; C.icnt[i - b0] = obmp::it_eof(it) ? obmp::G_CNT_LOOKUP : (uint16_t)0; }
        else if (valid && !m) {
            if (obmp::it_exact(it) || obmp::it_large(it)) { C.icnt[i - b0] = obmp::G_CNT_LOOKUP; any = true; }
            else C.icnt[i - b0] = (uint16_t)obmp::simple_count(it);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, m);
        if (m) C.mlist[n_ml + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = (uint8_t)(i - b0);
        n_ml += (uint32_t)__popc(bal);
    }
    __syncwarp();
    for (uint32_t k0 = 0; k0 < n_ml; k0 += 32) {
        const uint32_t k = k0 + lane; const bool on = k < n_ml;
        const bool staged = k0 == 0; /* the first W_MLCAP lines also stage their tuples */
        uint32_t ib = 0, d = 0; item_t it = 0;
        if (on) { ib = C.mlist[k]; it = A.items[X.i0 + b0 + ib]; d = X.doc_of(it); }
        obm_tuple *so = staged ? C.stage + k * obmp::W_LTS : nullptr; const uint32_t sc = staged ? obmp::W_LTS : 0u;
        uint32_t r;
        if constexpr (UNI) {
            const bool uni = on && obmp::it_unicode(it);
            r = k2_lex_lines(X, on && !uni, it, d, so, sc, nullptr, nullptr);
            if (__any_sync(0xffffffffu, uni)) r |= k2_lex_uni_lines(X, uni, it, d, so, sc, nullptr, nullptr);
        } else r = k2_lex_lines(X, on, it, d, so, sc, nullptr, nullptr);
        if (on) {
            C.icnt[ib] = (uint16_t)obmp::mres_tuples(r);
            if (obmp::mres_irregular(r)) { atomicOr(&A.doc_flag[d], obmp::GF_INTERACT); any = true; }
        }
    }
    __syncwarp();
    n_ml_out = n_ml;
    return any;
}

/* exact tuple counts of the unit's flagged documents -> counts[] (large documents were counted by k_exact_count) */
__device__ __forceinline__ void k2_count_flagged_docs(K2Ctx &X) {
    for (uint32_t q = X.lane; q < X.nd; q += 32) {
        const uint32_t d = X.d0 + q, f = X.A.doc_flag[d];
        if (f && !(f & obmp::GF_LARGE)) {
            const uint64_t o0 = X.A.doc_off[d];
            obm::SmallSink sink(nullptr, 0);
            obmp::doc_exact(X.T, X.A.bytes + o0, (uint32_t)(X.A.doc_off[d + 1] - o0), sink);
            X.A.counts[d] = sink.n_tuples;
        }
    }
    __syncwarp();
}

/* items of flagged documents: everything counts 0 except the closing item, which stands for the whole document */
__device__ __forceinline__ void k2_apply_flags(K2Ctx &X, uint32_t b0, uint32_t b1) {
    for (uint32_t i = b0 + X.lane; i < b1; i += 32) {
        const item_t it = X.A.items[X.i0 + i];
        if (X.A.doc_flag[X.doc_of(it)]) X.C.icnt[i - b0] = obmp::it_eof(it) ? obmp::G_CNT_LOOKUP : (uint16_t)0;
    }
    __syncwarp();
}

__device__ __forceinline__ uint64_t k2_block_total(K2Ctx &X, uint32_t b0, uint32_t b1) {
    uint64_t sum = 0;
    for (uint32_t i = b0 + X.lane; i < b1; i += 32) {
        uint32_t c = X.C.icnt[i - b0];
        if (c == obmp::G_CNT_LOOKUP) c = X.A.counts[X.doc_of(X.A.items[X.i0 + i])];
        sum += c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    return sum;
}

/* final positions of the block starting at `at0`: comment / EOF tuples in place, exact documents, unstaged
 * lines; then the staged marker tuples, a lane per tuple.  Returns the block's tuple count. */
template <bool UNI>
__device__ __forceinline__ uint64_t k2_write_block(K2Ctx &X, uint32_t b0, uint32_t b1, uint32_t n_ml, uint64_t at0, bool stable) {
    K2Warp &C = X.C; const PipeArgs &A = X.A; const uint32_t lane = X.lane;
    uint64_t run = 0; uint32_t mrun = 0;
    for (uint32_t c0 = b0; c0 < b1; c0 += 32) {
        const uint32_t i = c0 + lane; const bool valid = i < b1;
        const item_t it = valid ? A.items[X.i0 + i] : 0;
        const uint32_t d = valid ? X.doc_of(it) : 0u;
        const bool m = valid && obmp::it_marker(it) && !(stable && A.doc_flag[d]); /* same membership rule as the block's mlist */
        const bool eof = valid && obmp::it_eof(it);
        uint32_t c = valid ? C.icnt[i - b0] : 0u;
        const bool lookup = c == obmp::G_CNT_LOOKUP;
        if (lookup) c = A.counts[d];
        uint64_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
        const uint64_t at = at0 + run + incl - c;
        const uint32_t bal = __ballot_sync(0xffffffffu, m);
        bool relex = false;
        if (m) {
            const uint32_t k = mrun + (uint32_t)__popc(bal & ((1u << lane) - 1u));
            if (c) {
                if (k < obmp::W_MLCAP && c <= obmp::W_LTS) C.moff[k] = at;
                else relex = X.writing; /* not staged, or more tuples than a staging slot: lexed again, straight to its place */
            }
        } else if (eof) {
            A.tuple_off[d + 1] = at + c;
            if (!lookup) {
                if (X.writing && at < A.out_cap) A.out[at] = OBM_TUPLE(OBM_K_EOF, obmp::it_ls(it), 0);
                X.lexemes++;
            } else if (!obmp::it_large(it)) {
                X.exact++;
                if (X.writing) {
                    const uint64_t o0 = A.doc_off[d];
                    const uint64_t roomv = at < A.out_cap ? A.out_cap - at : 0;
                    obm::SmallSink sink(A.out + at, roomv > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)roomv);
                    const int st = obmp::doc_exact(X.T, A.bytes + o0, (uint32_t)(A.doc_off[d + 1] - o0), sink);
                    X.markers += sink.n_markers; X.lexemes += sink.n_lexemes; X.fatal += (st == obm::RUN_FATAL) ? 1u : 0u;
                }
            }
        } else if (valid && c) {
            if (X.writing) obmp::plain_write(it, A.out, at, A.out_cap);
            X.lexemes++;
        }
        if (__any_sync(0xffffffffu, relex)) {
            const uint64_t roomv = at < A.out_cap ? A.out_cap - at : 0;
            uint32_t mk = 0, lx = 0; /* locals: taking the context's address would push it to the stack */
            const uint32_t rc = roomv > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)roomv;
            if constexpr (UNI) {
                const bool uni = relex && obmp::it_unicode(it);
                k2_lex_lines(X, relex && !uni, it, d, A.out + at, rc, &mk, &lx);
                if (__any_sync(0xffffffffu, uni)) k2_lex_uni_lines(X, uni, it, d, A.out + at, rc, &mk, &lx);
            } else k2_lex_lines(X, relex, it, d, A.out + at, rc, &mk, &lx);
            X.markers += mk; X.lexemes += lx;
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
        mrun += (uint32_t)__popc(bal);
    }
    __syncwarp();
    if (X.writing) {
        const uint32_t ns = n_ml < obmp::W_MLCAP ? n_ml : obmp::W_MLCAP;
        for (uint32_t k = 0; k < ns; k++) {
            const uint64_t at = C.moff[k];
            if (at == ~0ull) continue;
            const uint32_t c = C.icnt[C.mlist[k]];
            const bool on = lane < c;
            const obm_tuple tup = on ? C.stage[k * obmp::W_LTS + lane] : 0;
            if (on && at + lane < A.out_cap) A.out[at + lane] = tup;
            const uint32_t kind = OBM_TUPLE_KIND(tup);
            const uint32_t mk = __ballot_sync(0xffffffffu, on && kind == OBM_K_MARKER_START);
            const uint32_t lx = __ballot_sync(0xffffffffu, on && (kind - (uint32_t)OBM_K_PART) > 4u);
            if (lane == 0) { X.markers += (uint32_t)__popc(mk); X.lexemes += (uint32_t)__popc(lx); }
        }
    }
    __syncwarp();
    return run;
}

/* UNI = false: no document of the batch needs per-line Unicode lexing (the common case; this instantiation carries no
 * Unicode line lexer, which keeps its register allocation and code footprint); UNI = true: the other batches.  Both are
 * launched, the one that does not apply returns at once. */
template <bool UNI>
__global__ void __launch_bounds__(obmp::W_WARPS * 32, 5)
k2_units(PipeArgs A) {
    if (A.ctl[CT_OVF] || (A.ctl[CT_UNI] != 0) != UNI) return; /* work records overflowed in k1: the host redoes the batch with the exact kernels */
    __shared__ K2Warp WS[obmp::W_WARPS];
    const obm::Tables T = dev_tables();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nunits = (uint32_t)A.ubase[A.ntiles];
    K2Ctx X{A, WS[threadIdx.x >> 5], T, 0, 0, 0, lane, A.out != nullptr && A.out_cap != 0, 0, 0, 0, 0};
    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(&A.ctl[CT_T2], 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= nunits) break;
        const obmp::Unit U = A.units[u];
        const uint32_t n_items = obmp::unit_items(U);
        X.i0 = U.item_base; X.d0 = U.doc_base; X.nd = obmp::unit_nd(U);
        uint64_t total; uint32_t n_ml = 0;
        const bool one_block = n_items <= obmp::W_ICAP;
        if (one_block) {
            const bool any = __any_sync(0xffffffffu, k2_lex_block<UNI>(X, 0, n_items, false, n_ml));
            if (any) { k2_count_flagged_docs(X); k2_apply_flags(X, 0, n_items); }
            total = k2_block_total(X, 0, n_items);
        } else {
            /* large unit: count sweep in 32-item blocks; only if it met flagged documents (K1 flags or interacting
             * lines found on the way) the flags are settled and the sweep is repeated; the write sweep lexes again */
            total = 0; bool any = false;
            for (uint32_t b0 = 0; b0 < n_items; b0 += 32) {
                const uint32_t b1 = min(b0 + 32u, n_items);
                uint32_t nm;
                any |= k2_lex_block<UNI>(X, b0, b1, false, nm);
                total += k2_block_total(X, b0, b1);
            }
            if (__any_sync(0xffffffffu, any)) {
                k2_count_flagged_docs(X); /* flags are final now: no line of an unflagged document is irregular */
                total = 0;
                for (uint32_t b0 = 0; b0 < n_items; b0 += 32) {
                    const uint32_t b1 = min(b0 + 32u, n_items);
                    uint32_t nm;
                    k2_lex_block<UNI>(X, b0, b1, true, nm);
                    total += k2_block_total(X, b0, b1);
                }
            }
        }
        const uint64_t base = obmf::lookback2_warp(A.st_tuples, A.st_blocks, u, nunits, total);
        if (lane == 0) {
            if (u == 0) A.tuple_off[0] = 0;
            if (u == nunits - 1 && A.out && base + total > A.out_cap) A.status[0] = 1;
        }
        if (one_block) k2_write_block<UNI>(X, 0, n_items, n_ml, base, false);
        else {
            uint64_t at = base;
            for (uint32_t b0 = 0; b0 < n_items; b0 += 32) {
                const uint32_t b1 = min(b0 + 32u, n_items);
                uint32_t nm;
                k2_lex_block<UNI>(X, b0, b1, true, nm);
                at += k2_write_block<UNI>(X, b0, b1, nm, at, true);
            }
        }
    }
    uint32_t markers = X.markers, lexemes = X.lexemes, exact = X.exact, fatal = X.fatal;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        markers += __shfl_down_sync(0xffffffffu, markers, o); lexemes += __shfl_down_sync(0xffffffffu, lexemes, o);
        exact += __shfl_down_sync(0xffffffffu, exact, o); fatal += __shfl_down_sync(0xffffffffu, fatal, o);
    }
    if (lane == 0) {
        if (markers) atomicAdd(&A.totals[0], (unsigned long long)markers);
        if (lexemes) atomicAdd(&A.totals[1], (unsigned long long)lexemes);
        if (exact) atomicAdd(&A.status[1], exact);
        if (fatal) atomicAdd(&A.status[2], fatal);
    }
}

} /* namespace obmq */
