/*
 * obm_fast.cuh -- the tile fast path: one persistent CTA per SM slot streams TILE-byte ranges of the
 * packed batch through shared memory (TMA bulk copy, cp.async.bulk + mbarrier), lexes every
 * document that starts in the range line-parallel (obm_tile.h) and writes the tuples at their final
 * position in ONE pass over the input: per-tile tuple counts are chained with a decoupled
 * look-back, so `doc_tuple_off` and the tuple stream come out ordered without a second read.
 *
 * HBM traffic per input byte: 1 B read (TMA) + ~0.32 B of tuples written + 8 B/document of offsets.
 * Documents larger than MAXDOC are handed to the exact kernels (k_exact_count / k_exact_fill in
 * obm_lib.cu) through a device-side list; their counts are folded into the same look-back chain.
 */
#pragma once
#include "obm_tile.h"

struct obm_handle;

namespace obmf {

using obmt::Smem;

/* ---- PTX helpers: mbarrier + 1-D TMA bulk copy global -> shared ---------------------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

/* ---- block-wide exclusive scan of one u32 per thread (NT threads); two barriers -------------------- */
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t *scratch /* NT/32+1 */, uint32_t &total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
    if (lane == 31) scratch[wid] = incl;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
#pragma unroll
    for (uint32_t w = 0; w < obmt::NT / 32; w++) { uint32_t s = scratch[w]; if (w < wid) pre += s; tot += s; }
    __syncthreads();
    total = tot;
    return pre + incl - v;
}

/* ---- look-back chain ----------------------------------------------------------------------------- */
#define LB_AGG (1ull << 62)
#define LB_INCL (2ull << 62)
#define LB_MASK ((1ull << 62) - 1)

/* Executed by warp 0 (all 32 lanes).  Publishes this tile's aggregate, then inspects 32 predecessors
 * per step (one coalesced load, two ballots) until an inclusive prefix is found.  Returns the
 * exclusive prefix (sum of all earlier tiles) in every lane. */
__device__ __forceinline__ uint64_t lookback_warp(volatile uint64_t *state, uint32_t t, uint64_t my_total) {
    const uint32_t lane = threadIdx.x & 31;
    if (t == 0) {
        if (lane == 0) { state[0] = LB_INCL | my_total; __threadfence(); }
        return 0;
    }
    if (lane == 0) { state[t] = LB_AGG | my_total; __threadfence(); }
    uint64_t sum = 0;
    int64_t hi = (int64_t)t - 1; /* nearest predecessor not yet accounted for */
    for (;;) {
        int64_t idx = hi - (int64_t)lane;
        uint64_t s = idx >= 0 ? state[idx] : LB_INCL; /* virtual inclusive 0 in front of tile 0 */
        uint32_t f = (uint32_t)(s >> 62);
        uint32_t incl = __ballot_sync(0xffffffffu, f == 2);
        uint32_t zero = __ballot_sync(0xffffffffu, f == 0);
        uint32_t upto = incl ? (uint32_t)__ffs((int)incl) - 1u : 31u;            /* last lane that contributes */
        uint32_t need = upto == 31u ? 0xffffffffu : ((2u << upto) - 1u);
        if (zero & need) { __nanosleep(40); continue; }                          /* a needed predecessor has not published yet */
        uint64_t v = lane <= upto ? (s & LB_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        sum += v;
        if (incl) break;
        hi -= 32;
    }
    if (lane == 0) { state[t] = LB_INCL | (sum + my_total); __threadfence(); }
    return sum;
}

/* Read-only walk (all 32 lanes of a warp): exclusive prefix of chain element t, i.e. the sum of all elements
 * before it, from whatever mixture of aggregates and inclusive prefixes their owners have published so far. */
__device__ __forceinline__ uint64_t lookback_read_warp(volatile uint64_t *state, uint32_t t) {
    const uint32_t lane = threadIdx.x & 31;
    if (t == 0) return 0;
    uint64_t sum = 0;
    int64_t hi = (int64_t)t - 1;
    for (;;) {
        int64_t idx = hi - (int64_t)lane;
        uint64_t s = idx >= 0 ? state[idx] : LB_INCL;
        uint32_t f = (uint32_t)(s >> 62);
        uint32_t incl = __ballot_sync(0xffffffffu, f == 2);
        uint32_t zero = __ballot_sync(0xffffffffu, f == 0);
        uint32_t upto = incl ? (uint32_t)__ffs((int)incl) - 1u : 31u;
        uint32_t need = upto == 31u ? 0xffffffffu : ((2u << upto) - 1u);
        if (zero & need) { __nanosleep(100); continue; }
        uint64_t v = lane <= upto ? (s & LB_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        sum += v;
        if (incl) break;
        hi -= 32;
    }
    return sum;
}

/* Two-level chain for many small elements (one per warp): level 0 holds plain aggregates, level 1 one
 * aggregate / inclusive prefix per block of 32 elements.  An element waits only for the aggregates of the
 * earlier elements of its own block and for a 32-wide walk over blocks, so one round trip advances the
 * resolved frontier by up to 1024 elements instead of 32.  Executed by all 32 lanes; returns the exclusive
 * prefix of element t (n = number of elements). */
__device__ __forceinline__ void lookback2_publish(volatile uint64_t *st0, uint32_t t, uint64_t my_total) {
    if ((threadIdx.x & 31) == 0) { st0[t] = LB_AGG | my_total; __threadfence(); }
}
/* the waiting half; may run long after lookback2_publish (the fused warp kernel scans its next unit in between) */
__device__ __forceinline__ uint64_t lookback2_resolve(volatile uint64_t *st0, volatile uint64_t *st1, uint32_t t, uint32_t n, uint64_t my_total) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t b = t >> 5, r = t & 31;
    const uint32_t bs = min(32u, n - (b << 5));
    uint64_t s;
    for (;;) {
        s = lane < r ? st0[(b << 5) + lane] : LB_AGG;
        if (__ballot_sync(0xffffffffu, (s >> 62) == 0) == 0) break;
        __nanosleep(100);
    }
    uint64_t partial = s & LB_MASK;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) partial += __shfl_xor_sync(0xffffffffu, partial, o);
    const uint64_t P = (r == bs - 1) ? lookback_warp(st1, b, partial + my_total) : lookback_read_warp(st1, b);
    return P + partial;
}
__device__ __forceinline__ uint64_t lookback2_warp(volatile uint64_t *st0, volatile uint64_t *st1, uint32_t t, uint32_t n, uint64_t my_total) {
    lookback2_publish(st0, t, my_total);
    return lookback2_resolve(st0, st1, t, n, my_total);
}

/* ---- pre-kernel: tile -> first document map, list of large documents ----------------------------- */
__global__ void __launch_bounds__(256)
k_tile_index(const uint64_t *__restrict__ doc_off, uint32_t ndocs, uint32_t ntiles, uint32_t *__restrict__ tile_first,
             uint32_t *__restrict__ large_list, uint32_t *__restrict__ n_large) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > ndocs) return;
    /* tiles t with  off[d-1] < t*TILE <= off[d]  have tile_first[t] = d; d == ndocs closes the table */
    uint64_t tprev_plus1 = d == 0 ? 0 : doc_off[d - 1] / obmt::TILE + 1;
    uint64_t tcur = d == ndocs ? (uint64_t)ntiles : doc_off[d] / obmt::TILE;
    if (d == ndocs && ndocs > 0 && tprev_plus1 > tcur) return;
    for (uint64_t t = tprev_plus1; t <= tcur && t <= ntiles; t++) tile_first[t] = d;
    if (d < ndocs && doc_off[d + 1] - doc_off[d] > obmt::MAXDOC) large_list[atomicAdd(n_large, 1u)] = d;
}

/* ---- the tile kernel ------------------------------------------------------------------------------- */
struct TileArgs {
    const uint8_t *bytes; const uint64_t *doc_off; uint32_t ndocs; uint64_t total_bytes;
    const uint32_t *tile_first; uint32_t ntiles;
    const uint32_t *counts;            /* precomputed tuple counts of large documents */
    obm_tuple *out; uint64_t out_cap; uint64_t *tuple_off;
    uint64_t *tile_state; uint32_t *ticket;
    uint32_t *status; unsigned long long *totals;
    uint32_t msplit;                   /* marker lines are dealt to this many warps (1..NT/32) */
};

struct CtaShared {
    Smem S;
    alignas(8) uint64_t mbar;
    uint64_t base;      /* look-back result */
    uint32_t tile;
    uint32_t sub_total; /* tuples of the current sub-batch */
    uint32_t stats[4];  /* markers, lexemes, exact docs, fatal docs */
};

/* Stages documents [da, db) (all small) and runs P1..P6.  Leaves S ready for the fill pass and
 * returns the sub-batch's tuple count (uniform across the CTA). */
__device__ __forceinline__ uint32_t subbatch_count(CtaShared &C, const TileArgs &A, const obm::Tables &T, uint32_t da, uint32_t db,
                                                   uint32_t &mbar_phase) {
    Smem &S = C.S;
    const uint32_t tid = threadIdx.x;
    const uint32_t nd = db - da;
    const uint64_t b0 = A.doc_off[da], b1 = A.doc_off[db];
    const uint64_t abs0 = (uint64_t)(uintptr_t)A.bytes + b0;
    const uint64_t base_abs = abs0 & ~15ull;
    const uint32_t skew = (uint32_t)(abs0 - base_abs);
    const uint32_t span = (uint32_t)(b1 - b0) + skew;   /* bytes of `data` in use */
    const uint32_t load = (span + 15u) & ~15u;
    /* P1: TMA bulk copy of the byte range into shared memory */
    if (tid == 0) {
        S.nd = nd; S.lo_pos = skew; S.hi_pos = span; S.n_owners = 0;
        if (load) {
            fence_proxy_async();
            mbar_expect_tx(&C.mbar, load);
            tma_bulk_g2s(S.data, (const void *)(uintptr_t)base_abs, load, &C.mbar);
        }
    }
    if (tid <= nd) S.dstart[tid] = (uint32_t)(A.doc_off[da + tid] - b0) + skew;
    __syncthreads();
    if (load) { mbar_wait(&C.mbar, mbar_phase); mbar_phase ^= 1; }
    /* P2: classify 32-byte words (strided: conflict-free shared loads, warp = 32 consecutive words) */
    const uint32_t nwords = (span + 31) >> 5;
    for (uint32_t base = 0; base < nwords; base += obmt::NT) {
        uint32_t wi = base + tid;
        if (wi < ((nwords + 31u) & ~31u) && wi < obmt::NW) obmt::classify_word(S, wi); /* whole warps participate (ballot) */
    }
    for (uint32_t wi = ((nwords + 31u) & ~31u) + tid; wi < obmt::NW; wi += obmt::NT) { S.nlw[wi] = 0; S.spw[wi] = 0; }
    __syncthreads();
    /* P3: per-document preparation */
    if (tid < nd) obmt::doc_prep(S, tid);
    __syncthreads();
    /* P4: bit-parallel line scan.  Thread t owns words [4t, 4t+4): first event of every line via carry
     * ripple, carries across threads/warps by generate/propagate look-ahead; then newline prefix counts
     * and the ordered list of lines that own tuples (position of their first special byte). */
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
        const uint32_t C = obmt::carry_lookahead32(Gb, Pb, cin, &dummy);
        obmt::first_events(nl, sp, lm, (C >> lane) & 1u, &lb);
    }
    uint32_t my_owners = 0, my_nl = 0;
#pragma unroll
    for (uint32_t j = 0; j < obmt::WPT; j++) { my_owners += (uint32_t)__popc(lb.own[j]); my_nl += (uint32_t)__popc(nl[j]); }
    uint32_t tot;
    uint32_t pre = block_scan_excl(my_nl | (my_owners << 16), S.scan_tmp, tot);
    {
        uint32_t nlp = pre & 0xFFFFu, own = pre >> 16;
#pragma unroll
        for (uint32_t j = 0; j < obmt::WPT; j++) { S.nlpre[tid * obmt::WPT + j] = (uint16_t)nlp; nlp += (uint32_t)__popc(nl[j]); }
        uint32_t n_owners = tot >> 16;
        if (n_owners <= obmt::QMAX) {
#pragma unroll
            for (uint32_t j = 0; j < obmt::WPT; j++) {
                uint32_t bits = lb.own[j];
                while (bits) { S.owner[own++] = (tid * obmt::WPT + j) * 32 + (uint32_t)(__ffs((int)bits) - 1); bits &= bits - 1; }
            }
        }
        if (tid == 0) { S.n_owners = n_owners <= obmt::QMAX ? n_owners : 0; S.n_markers_q = 0; }
        if (n_owners > obmt::QMAX && tid < nd) S.dflag[tid] |= obmt::DF_QOVERFLOW; /* too many special lines: exact path */
    }
    __syncthreads();
    /* P5a: every owner -> document, plain-line counts, dense list of marker lines */
    const uint32_t n_owners = S.n_owners;
    for (uint32_t o = tid; o < n_owners; o += obmt::NT) obmt::owner_prepare(S, o);
    __syncthreads();
    /* P5b: marker lines, dense: tokenize once, stage the tuples in shared memory */
    {
        /* marker line m runs on warp (m % msplit), lane (m / msplit) % 32: divergent lanes of one warp
         * serialise, so dealing the lines to several warps shortens this phase (the CTA's critical path) */
        const uint32_t nm = S.n_markers_q, ms = A.msplit;
        const uint32_t wid = tid >> 5, lane = tid & 31;
        if (wid < ms) for (uint32_t m = wid + lane * ms; m < nm; m += 32 * ms) obmt::marker_stage(S, T, m);
    }
    __syncthreads();
    /* P6: E[] = exclusive scan of owner counts (owners of irregular documents contribute nothing) */
    {
        uint32_t v[4], sum = 0;
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
            uint32_t o = tid * 4 + j;
            v[j] = (o < n_owners && !S.dflag[S.odoc[o] & 0x7Fu]) ? S.ocnt[o] : 0;
            sum += v[j];
        }
        uint32_t etot;
        uint32_t e = block_scan_excl(sum, S.scan_tmp, etot);
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) { uint32_t o = tid * 4 + j; if (o < obmt::QMAX) S.ocnt[o] = e; e += v[j]; }
        if (tid == 0) S.ocnt[obmt::QMAX] = etot;
    }
    if (tid <= nd) S.dfirst[tid] = tid == nd ? n_owners : obmt::first_owner_at(S, S.dstart[tid]);
    __syncthreads();
    if (n_owners < obmt::QMAX && tid == 0) S.ocnt[n_owners] = S.ocnt[obmt::QMAX]; /* E[n_owners] = total */
    __syncthreads();
    if (tid < nd) obmt::doc_count(S, T, tid);
    __syncthreads();
    {
        uint32_t v = tid < nd ? S.dcnt[tid] : 0, dtot;
        uint32_t e = block_scan_excl(v, S.scan_tmp, dtot);
        if (tid < nd) S.dcnt[tid] = e;
        if (tid == 0) { S.dcnt[nd] = dtot; C.sub_total = dtot; }
    }
    __syncthreads();
    return C.sub_total;
}

__device__ __forceinline__ void subbatch_fill(CtaShared &C, const TileArgs &A, const obm::Tables &T, uint32_t da, uint32_t db, uint64_t base) {
    Smem &S = C.S;
    const uint32_t tid = threadIdx.x, nd = db - da;
    obmt::FillStats fs = {0, 0, 0, 0};
    const uint32_t n_owners = S.n_owners;
    if (A.out) {
        for (uint32_t o = tid; o < n_owners; o += obmt::NT) obmt::owner_fill_thread(S, T, o, S.ocnt[o + 1] - S.ocnt[o], A.out, A.out_cap, base, fs);
        const uint32_t nm = S.n_markers_q < obmt::NSTAGE ? S.n_markers_q : obmt::NSTAGE;
        for (uint32_t m = tid >> 5; m < nm; m += obmt::NT / 32) obmt::owner_fill_staged(S, m, tid & 31, 32, A.out, A.out_cap, base, fs);
    }
    if (tid < nd) {
        A.tuple_off[da + tid] = base + S.dcnt[tid];
        if (A.out) obmt::doc_fill(S, T, tid, A.out, A.out_cap, base, fs);
    }
    /* stats: warp reduce, then shared atomics */
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        fs.markers += __shfl_down_sync(0xffffffffu, fs.markers, o); fs.lexemes += __shfl_down_sync(0xffffffffu, fs.lexemes, o);
        fs.exact_docs += __shfl_down_sync(0xffffffffu, fs.exact_docs, o); fs.fatal_docs += __shfl_down_sync(0xffffffffu, fs.fatal_docs, o);
    }
    if ((tid & 31) == 0) {
        if (fs.markers) atomicAdd(&C.stats[0], fs.markers);
        if (fs.lexemes) atomicAdd(&C.stats[1], fs.lexemes);
        if (fs.exact_docs) atomicAdd(&C.stats[2], fs.exact_docs);
        if (fs.fatal_docs) atomicAdd(&C.stats[3], fs.fatal_docs);
    }
}

__global__ void __launch_bounds__(obmt::NT, 3)
k_tile_scan(TileArgs A) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CtaShared &C = *reinterpret_cast<CtaShared *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    obm::Tables T;
    T.letter = D_GO_LETTER_RANGES; T.n_letter = D_GO_LETTER_RANGES_N;
    T.number = D_GO_NUMBER_RANGES; T.n_number = D_GO_NUMBER_RANGES_N;
    T.f64_overflow_digits = D_F64_OVERFLOW_DIGITS;
    if (tid == 0) { mbar_init(&C.mbar, 1); fence_mbar_init(); }
    if (tid < 4) C.stats[tid] = 0;
    __syncthreads();
    uint32_t mbar_phase = 0;
    for (;;) {
        if (tid == 0) C.tile = atomicAdd(A.ticket, 1u);
        __syncthreads();
        const uint32_t t = C.tile;
        if (t >= A.ntiles) break;
        const uint32_t d_first = A.tile_first[t], d_last = A.tile_first[t + 1];
        /* every document starting here is small except possibly the last one */
        uint32_t d_small_end = d_last;
        uint32_t large_cnt = 0;
        if (d_last > d_first && A.doc_off[d_last] - A.doc_off[d_last - 1] > obmt::MAXDOC) { d_small_end = d_last - 1; large_cnt = A.counts[d_last - 1]; }
        const bool single = d_small_end - d_first <= obmt::DMAX;
        uint64_t tile_total = large_cnt;
        for (uint32_t da = d_first; da < d_small_end; da += obmt::DMAX) {
            uint32_t db = min(da + obmt::DMAX, d_small_end);
            tile_total += subbatch_count(C, A, T, da, db, mbar_phase);
        }
        if (tid < 32) {
            uint64_t base = lookback_warp(A.tile_state, t, tile_total);
            if (tid == 0) {
                C.base = base;
                if (t == A.ntiles - 1) {
                    A.tuple_off[A.ndocs] = base + tile_total;
                    if (base + tile_total > A.out_cap) A.status[0] = 1;
                }
            }
        }
        __syncthreads();
        uint64_t base = C.base;
        for (uint32_t da = d_first; da < d_small_end; da += obmt::DMAX) {
            uint32_t db = min(da + obmt::DMAX, d_small_end);
            uint32_t sub = single ? C.sub_total : subbatch_count(C, A, T, da, db, mbar_phase);
            subbatch_fill(C, A, T, da, db, base);
            base += sub;
            __syncthreads();
        }
        if (d_small_end < d_last && tid == 0) A.tuple_off[d_last - 1] = base;
        __syncthreads();
    }
    if (tid < 4 && C.stats[tid]) {
        if (tid < 2) atomicAdd(&A.totals[tid], (unsigned long long)C.stats[tid]);
        else atomicAdd(&A.status[tid - 1], C.stats[tid]); /* status[1] = exact docs, status[2] = fatal docs */
    }
}

} /* namespace obmf */

/* scratch: tile_first u32[ntiles+2] | tile_state u64[ntiles+1] | large_list u32[max_large+1] | ctl u32[4] */
static inline uint64_t obm_fast_ntiles(uint64_t total_bytes) { return total_bytes / obmt::TILE + 1; }
/* table size for the list of large documents: a bound for every path's threshold (obmt::MAXDOC, obmw::MAXDOC > 8192) */
static inline uint64_t obm_fast_max_large(uint64_t total_bytes) { return total_bytes / 8192 + 1; }
static inline uint64_t obm_fast_scratch_bytes(uint32_t ndocs, uint64_t total_bytes) {
    (void)ndocs;
    uint64_t nt = obm_fast_ntiles(total_bytes);
    auto up = [](uint64_t v) { return (v + 255) / 256 * 256; };
    return up((nt + 2) * 4) + up((nt + 1) * 8) + up((obm_fast_max_large(total_bytes) + 1) * 4) + 256;
}
