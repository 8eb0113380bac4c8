/*
 * obm_pipe.cuh -- kernels of the three-stage pipeline (mode 0); logic in obm_pipe.h / obm_tile.h.
 *
 *   k1_scan     tile-resident classification + line logic -> items / marker-line records (no lexing)
 *   k2_markers  one thread per marker line, grid-stride over the device-side list: <WRITE=false> counts the
 *               line's tuples (and flags documents whose lines interact), <WRITE=true> runs after k3 and
 *               writes the tuples at their final position
 *   k3_doc_count / k3_doc_write   one thread per document: tuple counts (-> device-wide scan -> doc_tuple_off), then
 *               LINE/Comment of plain lines, EOF tuples, irregular documents (exact lexer), marker-line offsets
 */
#pragma once
#include "obm_fast.cuh"
#include "obm_pipe.h"

namespace obmq {

using obmt::SmemScan;
using obmp::item_t;
using obmp::MLine;

struct PipeArgs {
    const uint8_t *bytes; const uint64_t *doc_off; uint32_t ndocs; uint64_t total_bytes;
    const uint32_t *tile_first; uint32_t ntiles;
    /* K1 -> */
    item_t *items; uint32_t *item_slot; uint64_t items_cap;
    MLine *mlines; uint64_t mlines_cap;
    uint64_t *doc_item_off; uint32_t *doc_item_n; uint32_t *doc_flag;
    /* K2 -> */
    uint32_t *mres; uint64_t *moff;
    /* K3 -> */
    const uint32_t *counts; obm_tuple *out; uint64_t out_cap; uint64_t *tuple_off;
    uint64_t *tile_state; uint32_t *status; unsigned long long *totals;
    /* control words: [0] k1 ticket, [1] n_mlines, [2] k3 ticket, [3] items top (lo), [4] items top (hi), [5] overflow */
    uint32_t *ctl;
};
enum { CT_T1 = 0, CT_NML = 1, CT_T3 = 2, CT_ITOP = 4 /* u64 at ctl[4..5] */, CT_OVF = 6 };

__device__ __forceinline__ obm::Tables dev_tables() {
    obm::Tables T;
    T.letter = D_GO_LETTER_RANGES; T.n_letter = D_GO_LETTER_RANGES_N;
    T.number = D_GO_NUMBER_RANGES; T.n_number = D_GO_NUMBER_RANGES_N;
    T.f64_overflow_digits = D_F64_OVERFLOW_DIGITS;
    return T;
}

/* ---------------------------------------------------------------------------------------------- K1 -- */
struct K1Shared {
    SmemScan S;
    alignas(8) item_t sitems[obmt::QMAX]; /* items of the sub-batch (coalesced to HBM, searched per document) */
    alignas(8) uint64_t mbar;
    uint64_t item_base;
    uint32_t tile;
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
    for (;;) {
        if (tid == 0) C.tile = atomicAdd(&A.ctl[CT_T1], 1u);
        __syncthreads();
        const uint32_t t = C.tile;
        if (t >= A.ntiles) break;
        const uint32_t d_first = A.tile_first[t], d_last = A.tile_first[t + 1];
        uint32_t d_small_end = d_last;
        if (d_last > d_first && A.doc_off[d_last] - A.doc_off[d_last - 1] > obmt::MAXDOC) {
            d_small_end = d_last - 1;
            if (tid == 0) { A.doc_flag[d_last - 1] = obmp::GF_LARGE; A.doc_item_off[d_last - 1] = 0; A.doc_item_n[d_last - 1] = 0; }
        }
        for (uint32_t da = d_first; da < d_small_end; da += obmt::DMAX) {
            const uint32_t db = min(da + obmt::DMAX, d_small_end), nd = db - da;
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
            uint32_t n_owners = tot >> 16;
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
                if (tid == 0) {
                    unsigned long long top = atomicAdd(reinterpret_cast<unsigned long long *>(&A.ctl[CT_ITOP]), (unsigned long long)n_owners);
                    C.item_base = top;
                    if (top + n_owners > A.items_cap) A.ctl[CT_OVF] = 1;
                }
            }
            __syncthreads();
            /* P5 owners -> items (shared, then coalesced to HBM) + marker-line records */
            const uint64_t ibase = C.item_base;
            const bool room = ibase + n_owners <= A.items_cap;
            for (uint32_t o = tid; o < n_owners; o += obmt::NT) {
                obmp::K1Out r = obmp::k1_owner(S, o, da);
                uint32_t slot = 0xFFFFFFFFu;
                if (r.is_marker) {
                    slot = atomicAdd(&A.ctl[CT_NML], 1u);
                    if (slot < A.mlines_cap) A.mlines[slot] = r.ml; else A.ctl[CT_OVF] = 1;
                }
                sitems[o] = r.item;
                if (room) { A.items[ibase + o] = r.item; A.item_slot[ibase + o] = slot; }
            }
            __syncthreads();
            /* per document: first item, item count, flags */
            if (tid < nd) {
                /* items are in position order, hence grouped by document: binary search on the document field */
                uint32_t lo = 0, hi = n_owners;
                while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (obmp::it_doc(sitems[mid]) < tid) lo = mid + 1; else hi = mid; }
                uint32_t first = lo; hi = n_owners;
                while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (obmp::it_doc(sitems[mid]) <= tid) lo = mid + 1; else hi = mid; }
                A.doc_item_off[da + tid] = ibase + first;
                A.doc_item_n[da + tid] = lo - first;
                uint32_t f = S.dflag[tid];
                A.doc_flag[da + tid] = ((f & obmt::DF_NONASCII) ? obmp::GF_NONASCII : 0u) | ((f & obmt::DF_QOVERFLOW) ? obmp::GF_QOVERFLOW : 0u);
            }
            __syncthreads();
        }
    }
}

/* ---------------------------------------------------------------------------------------------- K2 -- */
template <bool WRITE>
__global__ void __launch_bounds__(256)
k2_markers(PipeArgs A) {
    if (A.ctl[CT_OVF]) return; /* work records overflowed in k1: the host redoes the batch with the exact kernels */
    const obm::Tables T = dev_tables();
    const uint32_t n = min(A.ctl[CT_NML], (uint32_t)min(A.mlines_cap, (uint64_t)0xFFFFFFFFu));
    uint32_t markers = 0, lexemes = 0;
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
        const MLine ml = A.mlines[m];
        const uint64_t o0 = A.doc_off[ml.doc];
        const uint32_t len = (uint32_t)(A.doc_off[ml.doc + 1] - o0);
        if (!WRITE) {
            uint32_t r = obmp::k2_marker_line(T, A.bytes + o0, len, ml, nullptr, 0);
            A.mres[m] = r;
            if (obmp::mres_irregular(r)) atomicOr(&A.doc_flag[ml.doc], obmp::GF_INTERACT);
        } else {
            if (A.doc_flag[ml.doc]) continue; /* the document went through the exact lexer in k3 */
            const uint64_t at = A.moff[m];
            const uint64_t roomv = at < A.out_cap ? A.out_cap - at : 0;
            uint32_t r = obmp::k2_marker_line(T, A.bytes + o0, len, ml, A.out + at, roomv > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)roomv, &markers, &lexemes);
            (void)r;
        }
    }
    if (WRITE) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { markers += __shfl_down_sync(0xffffffffu, markers, o); lexemes += __shfl_down_sync(0xffffffffu, lexemes, o); }
        if ((threadIdx.x & 31) == 0) {
            if (markers) atomicAdd(&A.totals[0], (unsigned long long)markers);
            if (lexemes) atomicAdd(&A.totals[1], (unsigned long long)lexemes);
        }
    }
}

/* ---------------------------------------------------------------------------------------------- K3 -- */
/* One thread per DOCUMENT, no shared memory, no barriers: a document's items are contiguous in HBM, so the
 * thread walks them twice -- k3_doc_count sums the tuple counts (the device-wide exclusive scan of those
 * counts, k_scan_* in obm_lib.cu, then yields doc_tuple_off), k3_doc_write materialises plain comment lines
 * and the EOF tuple and hands every marker line its final offset.  Flagged documents (non-ASCII, interacting
 * lines) run the exact lexer here; large documents keep the count k_exact_count wrote. */
__global__ void __launch_bounds__(256)
k3_doc_count(PipeArgs A, uint32_t *__restrict__ counts) {
    if (A.ctl[CT_OVF]) return; /* see k2_markers */
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= A.ndocs) return;
    const uint32_t flag = A.doc_flag[d];
    if (flag & obmp::GF_LARGE) return; /* counts[d] already holds the exact count */
    uint32_t c;
    if (flag) {
        const obm::Tables T = dev_tables();
        const uint64_t o0 = A.doc_off[d];
        obm::SmallSink sink(nullptr, 0);
        obmp::k3_doc_exact(T, A.bytes + o0, (uint32_t)(A.doc_off[d + 1] - o0), sink);
        c = sink.n_tuples;
    } else {
        const uint64_t i0 = A.doc_item_off[d];
        const uint32_t n = A.doc_item_n[d];
        c = 1; /* EOF */
        for (uint32_t i = 0; i < n; i++) {
            const item_t it = A.items[i0 + i];
            if (obmp::it_dead(it)) continue;
            c += obmp::it_marker(it) ? obmp::mres_tuples(A.mres[A.item_slot[i0 + i]]) : obmp::plain_count(it);
        }
    }
    counts[d] = c;
}

__global__ void __launch_bounds__(256)
k3_doc_write(PipeArgs A) {
    if (A.ctl[CT_OVF]) return;
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t markers = 0, lexemes = 0, exact = 0, fatal = 0;
    if (d < A.ndocs) {
        const uint32_t flag = A.doc_flag[d];
        const uint64_t base = A.tuple_off[d];
        const uint64_t o0 = A.doc_off[d];
        const uint32_t len = (uint32_t)(A.doc_off[d + 1] - o0);
        if (flag & obmp::GF_LARGE) {
            /* written by k_exact_fill */
        } else if (flag) {
            const obm::Tables T = dev_tables();
            const uint64_t roomv = (A.out && base < A.out_cap) ? A.out_cap - base : 0;
            obm::SmallSink sink(A.out + base, roomv > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)roomv);
            int st = obmp::k3_doc_exact(T, A.bytes + o0, len, sink);
            markers = sink.n_markers; lexemes = sink.n_lexemes; exact = 1; fatal = (st == obm::RUN_FATAL);
        } else {
            const uint64_t i0 = A.doc_item_off[d];
            const uint32_t n = A.doc_item_n[d];
            uint64_t at = base;
            for (uint32_t i = 0; i < n; i++) {
                const item_t it = A.items[i0 + i];
                if (obmp::it_dead(it)) continue;
                if (obmp::it_marker(it)) { const uint32_t slot = A.item_slot[i0 + i]; A.moff[slot] = at; at += obmp::mres_tuples(A.mres[slot]); }
                else { if (A.out) obmp::plain_write(it, A.out, at, A.out_cap); at += obmp::plain_count(it); lexemes++; }
            }
            if (A.out && at < A.out_cap) A.out[at] = OBM_TUPLE(OBM_K_EOF, len, 0);
            lexemes++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        markers += __shfl_down_sync(0xffffffffu, markers, o); lexemes += __shfl_down_sync(0xffffffffu, lexemes, o);
        exact += __shfl_down_sync(0xffffffffu, exact, o); fatal += __shfl_down_sync(0xffffffffu, fatal, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (markers) atomicAdd(&A.totals[0], (unsigned long long)markers);
        if (lexemes) atomicAdd(&A.totals[1], (unsigned long long)lexemes);
        if (exact) atomicAdd(&A.status[1], exact);
        if (fatal) atomicAdd(&A.status[2], fatal);
    }
}

} /* namespace obmq */
