/*
 * obm_group.cuh -- kernels of the ordered two-stage pipeline (mode 0); logic in obm_pipe.h / obm_tile.h.
 *
 *   k_tile_units  units (K1 sub-batches) per 16 KiB tile; an exclusive scan gives every unit a static id
 *   k1_scan       tile-resident classification + bit-parallel line logic (no lexing).  Emits ONE item stream
 *                 in global position order: a decoupled look-back over unit ids hands every unit its item
 *                 base, so the stream needs no per-document index.  Every document closes with an EOF item.
 *                 Units are cut into K2 groups of ~equal weight on the fly (second look-back chain).
 *   k2_group      one CTA per group (a few consecutive units = whole documents): marker items are compacted
 *                 and lexed ONCE, one thread per line, tuples staged in shared memory; item counts -> group
 *                 total -> decoupled look-back over groups -> final positions; comment / EOF tuples are
 *                 written in place, staged marker tuples copied out by warps.  Lines that do not fit the
 *                 staging area are lexed a second time straight to their final place; documents whose lines
 *                 interact (or that K1 flagged: non-ASCII, too many owning lines) are lexed by the exact
 *                 Unicode instantiation, all inside the group.  Writes doc_tuple_off as it goes.
 */
#pragma once
#include "obm_fast.cuh"
#include "obm_pipe.h"

namespace obmg {

using obmt::SmemScan;
using obmp::item_t;

struct GroupArgs {
    const uint8_t *bytes; const uint64_t *doc_off; uint32_t ndocs; uint64_t total_bytes;
    const uint32_t *tile_first; uint32_t ntiles;
    const uint64_t *ubase;      /* [ntiles+1] exclusive scan of units per tile; [ntiles] = number of units */
    /* K1 -> K2 */
    item_t *items; uint64_t items_cap;
    uint64_t *uitem;            /* [nunits+1] exclusive item prefix of the unit */
    uint32_t *udoc;             /* [nunits+1] first document of the unit */
    uint32_t *gstart; uint64_t gcap; /* first unit of group j */
    uint32_t *doc_flag;
    uint64_t *st_items, *st_weight, *st_tuples; /* look-back chains: units (items, weight), groups (tuples) */
    /* results */
    uint32_t *counts; obm_tuple *out; uint64_t out_cap; uint64_t *tuple_off;
    uint32_t *status; unsigned long long *totals;
    uint32_t *ctl;
};
enum { CT_T1 = 0, CT_T2 = 1, CT_NG = 2, CT_OVF = 6 };

__device__ __forceinline__ obm::Tables dev_tables() {
    obm::Tables T;
    T.letter = D_GO_LETTER_RANGES; T.n_letter = D_GO_LETTER_RANGES_N;
    T.number = D_GO_NUMBER_RANGES; T.n_number = D_GO_NUMBER_RANGES_N;
    T.f64_overflow_digits = D_F64_OVERFLOW_DIGITS;
    return T;
}

__global__ void __launch_bounds__(256)
k_tile_units(const uint64_t *__restrict__ doc_off, const uint32_t *__restrict__ tile_first, uint32_t ntiles, uint32_t *__restrict__ nsub) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint32_t d0 = tile_first[t], d1 = tile_first[t + 1];
    uint32_t n = 0;
    if (d1 > d0) {
        const uint32_t large = (doc_off[d1] - doc_off[d1 - 1] > obmt::MAXDOC) ? 1u : 0u;
        const uint32_t ns = d1 - d0 - large;
        n = (ns + obmt::DMAX - 1) / obmt::DMAX;
        if (n == 0) n = 1;
    }
    nsub[t] = n;
}

/* ---------------------------------------------------------------------------------------------- K1 -- */
struct K1Shared {
    SmemScan S;
    alignas(8) item_t sitems[obmt::QMAX]; /* items of the sub-batch in owner order */
    uint16_t dlast[obmt::DMAX + 1];       /* index after the last owner of document k */
    alignas(8) uint64_t mbar;
    uint64_t item_base, w_base;
    uint32_t tile, n_ml;
};

__global__ void __launch_bounds__(obmt::NT, 4)
k1_scan(GroupArgs A) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    K1Shared &C = *reinterpret_cast<K1Shared *>(smem_raw);
    SmemScan &S = C.S;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { obmf::mbar_init(&C.mbar, 1); obmf::fence_mbar_init(); }
    __syncthreads();
    uint32_t mbar_phase = 0;
    item_t *sitems = C.sitems;
    const uint64_t nunits = A.ubase[A.ntiles];
    for (;;) {
        if (tid == 0) C.tile = atomicAdd(&A.ctl[CT_T1], 1u);
        __syncthreads();
        const uint32_t t = C.tile;
        if (t >= A.ntiles) break;
        const uint32_t d_first = A.tile_first[t], d_last = A.tile_first[t + 1];
        if (d_last == d_first) { __syncthreads(); continue; }
        const bool has_large = A.doc_off[d_last] - A.doc_off[d_last - 1] > obmt::MAXDOC;
        const uint32_t d_small_end = d_last - (has_large ? 1u : 0u);
        uint32_t nsub = (d_small_end - d_first + obmt::DMAX - 1) / obmt::DMAX;
        if (nsub == 0) nsub = 1;
        const uint64_t u0 = A.ubase[t];
        if (has_large && tid == 0) A.doc_flag[d_last - 1] = obmp::GF_LARGE;
        for (uint32_t k = 0; k < nsub; k++) {
            const uint32_t da = d_first + k * obmt::DMAX, db = min(da + obmt::DMAX, d_small_end), nd = db - da;
            const uint32_t extra = (k == nsub - 1 && has_large) ? 1u : 0u;
            const uint64_t u = u0 + k;
            uint32_t n_owners = 0;
            if (tid == 0) C.n_ml = 0;
            if (nd) {
                const uint64_t b0 = A.doc_off[da], b1 = A.doc_off[db];
                const uint64_t abs0 = (uint64_t)(uintptr_t)A.bytes + b0, base_abs = abs0 & ~15ull;
                const uint32_t skew = (uint32_t)(abs0 - base_abs), span = (uint32_t)(b1 - b0) + skew, load = (span + 15u) & ~15u;
                if (tid == 0) {
                    S.nd = nd; S.lo_pos = skew; S.hi_pos = span; S.n_owners = 0;
                    if (load) { obmf::fence_proxy_async(); obmf::mbar_expect_tx(&C.mbar, load); obmf::tma_bulk_g2s(S.data, (const void *)(uintptr_t)base_abs, load, &C.mbar); }
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
                if (tid < nd) obmt::doc_prep(S, tid);
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
                /* P5 owners -> items (shared memory, owner order) */
                uint32_t my_ml = 0;
                for (uint32_t o = tid; o < n_owners; o += obmt::NT) {
                    const item_t it = obmp::k1_owner_item(S, o);
                    sitems[o] = it;
                    my_ml += obmp::it_marker(it) ? 1u : 0u;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) my_ml += __shfl_down_sync(0xffffffffu, my_ml, o);
                if ((tid & 31) == 0 && my_ml) atomicAdd(&C.n_ml, my_ml);
                __syncthreads();
                /* per document: index after its last owner (owners are in position order, hence grouped by document) */
                if (tid < nd) {
                    uint32_t lo = 0, hi = n_owners;
                    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (obmp::it_doc(sitems[mid]) <= tid) lo = mid + 1; else hi = mid; }
                    C.dlast[tid] = (uint16_t)lo;
                    const uint32_t f = S.dflag[tid];
                    A.doc_flag[da + tid] = ((f & obmt::DF_NONASCII) ? obmp::GF_NONASCII : 0u) | ((f & obmt::DF_QOVERFLOW) ? obmp::GF_QOVERFLOW : 0u);
                }
            }
            __syncthreads();
            /* ordered allocation: exclusive prefixes of items and weight over all earlier units */
            const uint32_t n_items = n_owners + nd + extra, n_ml = C.n_ml;
            const uint32_t weight = obmp::unit_weight(n_items, n_ml);
            if (tid < 32) {
                const uint64_t b = obmf::lookback_warp(A.st_items, (uint32_t)u, n_items);
                if (tid == 0) C.item_base = b;
            } else if (tid < 64) {
                const uint64_t b = obmf::lookback_warp(A.st_weight, (uint32_t)u, weight);
                if (tid == 32) C.w_base = b;
            }
            __syncthreads();
            const uint64_t ibase = C.item_base;
            const bool room = ibase + n_items <= A.items_cap;
            if (room) {
                for (uint32_t o = tid; o < n_owners; o += obmt::NT) { const item_t it = sitems[o]; A.items[ibase + o + obmp::it_doc(it)] = it; }
                if (tid < nd) A.items[ibase + C.dlast[tid] + tid] = obmp::make_eof_item(S.dstart[tid + 1] - S.dstart[tid], tid, S.dflag[tid] != 0);
                if (extra && tid == 0) A.items[ibase + n_items - 1] = obmp::make_large_item();
            }
            if (tid == 0) {
                if (!room) A.ctl[CT_OVF] = 1;
                A.uitem[u] = ibase; A.udoc[u] = da;
                const uint64_t w0 = C.w_base, w1 = w0 + weight;
                for (uint64_t j = w0 / obmp::GROUP_W + 1; j <= w1 / obmp::GROUP_W; j++) { if (j < A.gcap) A.gstart[j] = (uint32_t)(u + 1); else A.ctl[CT_OVF] = 1; }
                if (u == nunits - 1) { A.uitem[nunits] = ibase + n_items; A.udoc[nunits] = A.ndocs; A.ctl[CT_NG] = (uint32_t)(w1 / obmp::GROUP_W + 1); }
            }
            __syncthreads();
        }
    }
}

/* ---------------------------------------------------------------------------------------------- K2 -- */
struct K2Shared {
    alignas(16) obm_tuple stage[obmp::G_MLCAP * obmp::G_LTS];
    uint64_t moff[obmp::G_MLCAP];       /* staged line -> final output position (~0: not copied) */
    uint16_t icnt[obmp::G_IMAX];        /* tuples per item (G_CNT_LOOKUP: counts[doc]) */
    uint16_t mlist[obmp::G_IMAX];       /* marker rank -> item index */
    uint32_t ut_item[obmp::G_UCAP + 1]; /* unit -> first item (group-relative) */
    uint32_t ut_doc[obmp::G_UCAP + 1];  /* unit -> first document */
    uint64_t red[obmp::G_NT / 32 + 1];
    uint32_t red2[obmp::G_NT / 32 + 1];
    uint64_t base;
    uint32_t group, any_flag;
};

/* exclusive scan over the CTA of (v, f) with totals; v 64-bit, f small */
__device__ __forceinline__ void group_scan(K2Shared &C, uint64_t v, uint32_t f, uint64_t &v_excl, uint32_t &f_excl, uint64_t &v_tot, uint32_t &f_tot) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t vi = v; uint32_t fi = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t tv = __shfl_up_sync(0xffffffffu, vi, o); const uint32_t tf = __shfl_up_sync(0xffffffffu, fi, o);
        if (lane >= (uint32_t)o) { vi += tv; fi += tf; }
    }
    if (lane == 31) { C.red[wid] = vi; C.red2[wid] = fi; }
    __syncthreads();
    uint64_t pv = 0, tv = 0; uint32_t pf = 0, tf = 0;
#pragma unroll
    for (uint32_t w = 0; w < obmp::G_NT / 32; w++) { const uint64_t a = C.red[w]; const uint32_t b = C.red2[w]; if (w < wid) { pv += a; pf += b; } tv += a; tf += b; }
    __syncthreads();
    v_excl = pv + vi - v; f_excl = pf + fi - f; v_tot = tv; f_tot = tf;
}

__device__ __forceinline__ uint32_t unit_of(const K2Shared &C, uint32_t nu, uint32_t i) {
    uint32_t lo = 0, hi = nu; /* last k with ut_item[k] <= i */
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (C.ut_item[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}
__device__ __forceinline__ uint32_t doc_of_item(const K2Shared &C, uint32_t nu, uint32_t i, item_t it) {
    const uint32_t u = unit_of(C, nu, i);
    return obmp::it_large(it) ? C.ut_doc[u + 1] - 1u : C.ut_doc[u] + obmp::it_doc(it);
}

__global__ void __launch_bounds__(obmp::G_NT)
k2_group(GroupArgs A) {
    if (A.ctl[CT_OVF]) return; /* work records overflowed in k1: the host redoes the batch with the exact kernels */
    __shared__ K2Shared C;
    const obm::Tables T = dev_tables();
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t NG = A.ctl[CT_NG];
    const uint64_t nunits = A.ubase[A.ntiles];
    const bool writing = A.out != nullptr && A.out_cap != 0;
    uint32_t markers = 0, lexemes = 0, exact = 0, fatal = 0;
    for (;;) {
        if (tid == 0) C.group = atomicAdd(&A.ctl[CT_T2], 1u);
        __syncthreads();
        const uint32_t j = C.group;
        if (j >= NG) break;
        const uint64_t ua = A.gstart[j], ub = (j + 1 < NG) ? A.gstart[j + 1] : nunits;
        const uint32_t nu = (uint32_t)(ub - ua);
        const uint64_t i0 = A.uitem[ua];
        const uint32_t n_items = (uint32_t)(A.uitem[ub] - i0);
        for (uint32_t k = tid; k <= nu; k += obmp::G_NT) { C.ut_item[k] = (uint32_t)(A.uitem[ua + k] - i0); C.ut_doc[k] = A.udoc[ua + k]; }
        if (tid == 0) C.any_flag = 0;
        C.moff[tid] = ~0ull; /* G_MLCAP == G_NT */
        __syncthreads();
        /* pass A: compact the marker items, counts of everything else */
        uint32_t n_ml = 0;
        for (uint32_t c0 = 0; c0 < n_items; c0 += obmp::G_NT) {
            const uint32_t i = c0 + tid; const bool valid = i < n_items;
            const item_t it = valid ? A.items[i0 + i] : 0;
            const bool m = valid && obmp::it_marker(it);
            uint64_t ve, vt; uint32_t fe, ft;
            group_scan(C, 0, m ? 1u : 0u, ve, fe, vt, ft);
            if (m) C.mlist[n_ml + fe] = (uint16_t)i;
            else if (valid) {
                if (obmp::it_exact(it) || obmp::it_large(it)) { C.icnt[i] = obmp::G_CNT_LOOKUP; C.any_flag = 1; }
                else C.icnt[i] = (uint16_t)obmp::simple_count(it);
            }
            n_ml += ft;
        }
        __syncthreads();
        /* lex every marker line once; the first G_MLCAP lines stage their tuples in shared memory */
        for (uint32_t k = tid; k < n_ml; k += obmp::G_NT) {
            const uint32_t i = C.mlist[k];
            const item_t it = A.items[i0 + i];
            const uint32_t d = doc_of_item(C, nu, i, it);
            const uint64_t o0 = A.doc_off[d];
            const uint32_t len = (uint32_t)(A.doc_off[d + 1] - o0);
            const bool staged = k < obmp::G_MLCAP;
            const uint32_t r = obmp::k2_marker_item(T, A.bytes + o0, len, it, staged ? C.stage + k * obmp::G_LTS : nullptr, staged ? obmp::G_LTS : 0u);
            C.icnt[i] = (uint16_t)obmp::mres_tuples(r);
            if (obmp::mres_irregular(r)) { atomicOr(&A.doc_flag[d], obmp::GF_INTERACT); C.any_flag = 1; }
        }
        __syncthreads();
        const bool any = C.any_flag != 0;
        if (any) {
            /* documents that need the exact lexer: count them here (large documents were counted by k_exact_count) */
            const uint32_t dA = C.ut_doc[0], dB = C.ut_doc[nu];
            for (uint32_t d = dA + tid; d < dB; d += obmp::G_NT) {
                const uint32_t f = A.doc_flag[d];
                if (f && !(f & obmp::GF_LARGE)) {
                    const uint64_t o0 = A.doc_off[d];
                    obm::SmallSink sink(nullptr, 0);
                    obmp::k3_doc_exact(T, A.bytes + o0, (uint32_t)(A.doc_off[d + 1] - o0), sink);
                    A.counts[d] = sink.n_tuples;
                }
            }
            __syncthreads();
            for (uint32_t i = tid; i < n_items; i += obmp::G_NT) {
                const item_t it = A.items[i0 + i];
                if (A.doc_flag[doc_of_item(C, nu, i, it)]) C.icnt[i] = obmp::it_eof(it) ? obmp::G_CNT_LOOKUP : (uint16_t)0;
            }
            __syncthreads();
        }
        /* group total -> look-back -> base */
        uint64_t sum = 0;
        for (uint32_t i = tid; i < n_items; i += obmp::G_NT) {
            uint32_t c = C.icnt[i];
            if (c == obmp::G_CNT_LOOKUP) c = A.counts[doc_of_item(C, nu, i, A.items[i0 + i])];
            sum += c;
        }
        uint64_t se, gtotal; uint32_t fe0, ft0;
        group_scan(C, sum, 0, se, fe0, gtotal, ft0);
        if (tid < 32) {
            const uint64_t b = obmf::lookback_warp(A.st_tuples, j, gtotal);
            if (tid == 0) {
                C.base = b;
                if (j == 0) A.tuple_off[0] = 0;
                if (j == NG - 1) { A.tuple_off[A.ndocs] = b + gtotal; if (A.out && b + gtotal > A.out_cap) A.status[0] = 1; }
            }
        }
        __syncthreads();
        const uint64_t base = C.base;
        /* pass B: final positions; comment / EOF tuples in place, exact documents, unstaged lines */
        uint64_t run = 0; uint32_t mrun = 0;
        for (uint32_t c0 = 0; c0 < n_items; c0 += obmp::G_NT) {
            const uint32_t i = c0 + tid; const bool valid = i < n_items;
            const item_t it = valid ? A.items[i0 + i] : 0;
            const bool m = valid && obmp::it_marker(it);
            const bool eof = valid && obmp::it_eof(it);
            uint32_t c = valid ? C.icnt[i] : 0u;
            const bool lookup = c == obmp::G_CNT_LOOKUP;
            uint32_t d = 0;
            if (eof || lookup) d = doc_of_item(C, nu, i, it);
            if (lookup) c = A.counts[d];
            uint64_t ve, vt; uint32_t fe, ft;
            group_scan(C, c, m ? 1u : 0u, ve, fe, vt, ft);
            const uint64_t at = base + run + ve;
            if (m) {
                const uint32_t k = mrun + fe;
                if (c) {
                    if (k < obmp::G_MLCAP && c <= obmp::G_LTS) C.moff[k] = at;
                    else if (writing) {
                        const uint32_t dd = doc_of_item(C, nu, i, it);
                        const uint64_t o0 = A.doc_off[dd];
                        const uint64_t roomv = at < A.out_cap ? A.out_cap - at : 0;
                        obmp::k2_marker_item(T, A.bytes + o0, (uint32_t)(A.doc_off[dd + 1] - o0), it, A.out + at,
                                             roomv > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)roomv, &markers, &lexemes);
                    }
                }
            } else if (eof) {
                A.tuple_off[d + 1] = at + c;
                if (!lookup) {
                    if (writing && at < A.out_cap) A.out[at] = OBM_TUPLE(OBM_K_EOF, obmp::it_ls(it), 0);
                    lexemes++;
                } else if (!obmp::it_large(it)) {
                    const uint64_t o0 = A.doc_off[d];
                    const uint64_t roomv = (writing && at < A.out_cap) ? A.out_cap - at : 0;
                    obm::SmallSink sink(A.out + at, roomv > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)roomv);
                    const int st = obmp::k3_doc_exact(T, A.bytes + o0, (uint32_t)(A.doc_off[d + 1] - o0), sink);
                    markers += sink.n_markers; lexemes += sink.n_lexemes; exact++; fatal += (st == obm::RUN_FATAL) ? 1u : 0u;
                }
            } else if (valid && c) {
                if (writing) obmp::plain_write(it, A.out, at, A.out_cap);
                lexemes++;
            }
            run += vt; mrun += ft;
        }
        __syncthreads();
        /* staged marker tuples -> final position, a warp per line */
        if (writing) {
            const uint32_t ns = n_ml < obmp::G_MLCAP ? n_ml : obmp::G_MLCAP;
            for (uint32_t k = wid; k < ns; k += obmp::G_NT / 32) {
                const uint64_t at = C.moff[k];
                if (at == ~0ull) continue;
                const uint32_t c = C.icnt[C.mlist[k]];
                const bool on = lane < c;
                const obm_tuple tup = on ? C.stage[k * obmp::G_LTS + lane] : 0;
                if (on && at + lane < A.out_cap) A.out[at + lane] = tup;
                const uint32_t kind = OBM_TUPLE_KIND(tup);
                const uint32_t mk = __ballot_sync(0xffffffffu, on && kind == OBM_K_MARKER_START);
                const uint32_t lx = __ballot_sync(0xffffffffu, on && (kind - (uint32_t)OBM_K_PART) > 4u);
                if (lane == 0) { markers += (uint32_t)__popc(mk); lexemes += (uint32_t)__popc(lx); }
            }
        }
        __syncthreads();
    }
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

} /* namespace obmg */
