/*
 * Hand-written sm_100a scan kernels for the exact-string-matching hot path.
 *
 * What every kernel computes is the result of src/algos/bf.c:33-36 (reference tree): the
 * number of start positions j in the shard's owned range with x[0..m) == y[j..j+m), and,
 * when asked, those positions.  How they get there is B200-first:
 *
 *   - the text is streamed from HBM exactly once with coalesced, L1-bypassing 256-bit loads
 *     (ld.global.nc.L1::no_allocate.v8.u32 -> LDG.E.NA.ENL2.256), U independent loads in flight per
 *     thread, one warp per U KiB group, groups strided across a grid sized to the resident CTAs;
 *   - a packed candidate filter runs entirely in registers:
 *       * scan_block<B,L> (m >= L+B-1): every occurrence contains the L bytes that start at some
 *         B-aligned text position, so only ALIGNED windows are fingerprinted (IMAD, FMA pipe) and
 *         compared with the B fingerprints of the pattern's L-grams at offsets a..a+B-1 (one
 *         ISETP.EQ.OR per start position, ALU pipe): no shifts, no shared memory, no halo when L == B;
 *       * scan_window<NW,OFF> (5 <= m, low-entropy texts or m < 2B-1): an 8/12/16-byte window
 *         fingerprint at EVERY start position (SHF funnel shift + IMAD + ISETP.EQ.OR per position);
 *         the halo comes from the neighbouring lane by shuffle, not from memory;
 *       * scan_block<2,L> (5 <= m <= 10): the same with windows at every EVEN position (one funnel shift per word);
 *       * scan_swar<M> (m <= 4): exact byte-parallel compare, four start positions per instruction,
 *         counted with POPC in registers;
 *   - sub-linear variants read less than the text: scan_skip<G> samples one gram per `stride` bytes (skip
 *     families, long patterns), scan_packed<K> / scan_packed_half<KH> stream a 4/2/1-bit recoding of a text over
 *     <= 16 byte values; both resolve hits through a shared-memory bitmap and an exact table;
 *   - scan_multi<B> counts a whole pattern set in one pass; scan_block_tma<B> is the TMA-fed (cp.async.bulk +
 *     mbarrier ring) variant of scan_block kept as a measured experiment;
 *   - candidates (rare) go to an exact verification that reads the text through L2 with unaligned
 *     32-bit windows (warp-cooperative, 128 bytes per step, in scan_block);
 *   - counts are reduced with warp shuffles and ONE atomicAdd per block.
 *
 * Nothing here is a dense contraction: tensor cores / TMEM are deliberately unused; the bound
 * is HBM bandwidth, with integer issue rate as the secondary limit.
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sb {

constexpr int kThreads = 256;       // threads per CTA
#ifndef SB_MINB
#define SB_MINB 1                   // __launch_bounds__ min CTAs/SM for the streaming kernels (tuning knob)
#endif
constexpr int kPadBytes = 8192;     // every text buffer is zero padded to a multiple of this plus this much
constexpr unsigned kFull = 0xffffffffu;

// fingerprint multipliers (odd => w -> w*K is a bijection mod 2^32)
constexpr uint32_t K1 = 0x9E3779B1u;
constexpr uint32_t K2 = 0x85EBCA77u;
constexpr uint32_t K3 = 0xC2B2AE3Du;

__host__ __device__ __forceinline__ uint32_t fp8(uint32_t w0, uint32_t w1) { return w0 * K1 + w1; }
__host__ __device__ __forceinline__ uint32_t fp16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
    return fp8(w0, w1) * K2 + fp8(w2, w3);
}
__host__ __device__ __forceinline__ uint32_t fp32(uint32_t lo16, uint32_t hi16) { return lo16 * K3 + hi16; }

struct ScanArgs {
    const uint8_t *text;          // shard buffer, 32-B aligned, zero padded (kPadBytes rule)
    const uint32_t *packed;       // scan_packed: 2-bit (1-bit for two symbols) recoding of the shard, or nullptr
    int cpw;                      // scan_packed: characters per packed word (16 or 32)
    const uint32_t *pat;          // pattern as little-endian words, zero padded
    unsigned long long *counters; // [0] match count, [2] verified candidates
    uint32_t *pos;                // position output: bitmap over the scanned start positions (bit j <=> an occurrence
                                  // starts at local position j), or nullptr; compacted in order by bitmap_write_kernel
    unsigned long long base;      // global offset of local byte 0
    long long hi;                 // owned local start positions are [0, hi)
    unsigned long long ngroups;   // groups to scan (group size depends on the kernel)
    int m;
    int anchor;                   // byte offset a of the fingerprinted sub-pattern
    int nwords;                   // ceil(m/4)
    uint32_t tailmask;            // mask for the last pattern word
    uint32_t H[32];               // scan_block: fingerprints of x[a+d .. a+d+B), d < B
    uint32_t X0, X1, M0;          // scan_swar: masked first word; scan_window: first and (NW=2) second word
    uint32_t Xw[3];               // scan_window: window words 1..NW-1 (Xw[NW-2] is the last one)
    uint32_t G;                   // scan_window: fingerprint of the window words
    // scan_skip: one G-byte gram is sampled every `stride` bytes; aux = Bloom bitmap (kBloomWords words)
    // followed by an open-addressing table {fingerprint of x[anchor+d .. anchor+d+G), d+1}, d < stride
    const uint32_t *aux;
    unsigned long long nsamples;
    int stride;
    // scan_multi: npat patterns of the same length, pattern p at pat + p*pat_stride_words; per-pattern
    // counts at multi_counts[p]; aux = Bloom bitmap + table {fingerprint, ((p << 5) | d) + 1}
    unsigned long long *multi_counts;
    int npat, pat_stride_words;
    // smallest period of the pattern, m - mpNext[m] (mp.c:26-35), when it is at most m/2; else 0.
    // Drives period-aware verification in scan_block: after a verified occurrence at j, a candidate at
    // j + k*period (< j + m) only needs its last k*period bytes compared.
    int period;
    // word index of the pattern word that holds its RAREST byte (by the text's byte histogram): one of the three
    // words the lane-parallel pre-check of a candidate looks at (first, last, rarest) before the exact verification;
    // -1: no byte of the pattern is rare, two probe words
    int probe;
};

constexpr int kBloomBits = 18;                       // 2^18-bit Bloom bitmap = 32 KiB of shared memory, 2 probes
constexpr int kBloomWords = (1 << kBloomBits) / 32;
__host__ __device__ __forceinline__ uint32_t bloom_slot(uint32_t h) { return (h * 0x2545F491u) >> (32 - kBloomBits); }
__host__ __device__ __forceinline__ uint32_t bloom_slot2(uint32_t h) { return (h * 0x9FB21C65u + 0x7F4A7C15u) >> (32 - kBloomBits); }
__host__ __device__ __forceinline__ uint32_t bloom_slot3(uint32_t h) { return ((h ^ (h >> 15)) * 0x2C1B3C6Du) >> (32 - kBloomBits); }
// exact second level: open-addressing table {fingerprint, d+1} (linear probing, 0 = empty) after the bitmap
constexpr int kSkipTableBits = 13;
constexpr int kSkipTableSize = 1 << kSkipTableBits;   // 8192 entries >= 2 x 4193 grams
__host__ __device__ __forceinline__ uint32_t table_slot(uint32_t h) { return (h * 0xB5297A4Du) >> (32 - kSkipTableBits); }

struct V8 {
    uint32_t w[8];
};

#ifndef SB_STREAM_L2
#define SB_STREAM_L2 "" /* optional L2 prefetch-size hint for the streaming loads (tuning) */
#endif
__device__ __forceinline__ V8 ldg_stream256(const uint8_t *p)
{
    V8 r;
    asm volatile("ld.global.nc.L1::no_allocate" SB_STREAM_L2 ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint4 ldg_stream128(const uint8_t *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

/* 32-bit little-endian window at an arbitrary byte offset (two aligned loads + funnel shift) */
__device__ __forceinline__ uint32_t text_window32(const uint8_t *text, long long off)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(text) + (off >> 2);
    uint32_t a = __ldg(w), b = __ldg(w + 1);
    return __funnelshift_r(a, b, (unsigned)(off & 3) * 8u);
}

/* Exact comparison of x[0..m) with text[j..j+m), one 32-bit window per lane per step.
 * All 32 lanes must call it with the same j.  Words [from_word, nwords) are compared; for long patterns the
 * first step looks at BOTH ends (16 words each): a candidate that agrees with the pattern around the
 * fingerprinted block but differs far from it (near-periodic inputs) is then rejected in one step. */
__device__ __forceinline__ bool warp_compare(const ScanArgs &a, long long j, const uint32_t *pat, int k, bool active)
{
    bool ok = true;
    if (active) {
        uint32_t diff = text_window32(a.text, j + 4ll * k) ^ __ldg(pat + k);
        if (k == a.nwords - 1) diff &= a.tailmask;
        ok = (diff == 0);
    }
    return __all_sync(kFull, ok);
}

__device__ __forceinline__ bool warp_verify(const ScanArgs &a, long long j, unsigned lane, const uint32_t *pat = nullptr,
                                            int from_word = 0)
{
    const int nwords = a.nwords;
    if (!pat) pat = a.pat;
    int lo = from_word, hi = nwords;
    if (hi - lo > 32) {
        const int k = lane < 16 ? lo + (int)lane : hi - 32 + (int)lane;
        if (!warp_compare(a, j, pat, k, true)) return false;
        lo += 16;
        hi -= 16;
    }
    for (int base = lo; base < hi; base += 32) {
        const int k = base + (int)lane;
        if (!warp_compare(a, j, pat, k, k < hi)) return false;
    }
    return true;
}

/* the same comparison by a single lane (divergent callers), words [from, nwords) */
__device__ __forceinline__ bool lane_verify(const ScanArgs &a, long long j, int from)
{
    for (int k = from; k < a.nwords; k++) {
        uint32_t diff = text_window32(a.text, j + 4ll * k) ^ __ldg(a.pat + k);
        if (k == a.nwords - 1) diff &= a.tailmask;
        if (diff) return false;
    }
    return true;
}

/* Pre-check of ONE candidate per lane: three probe words - the first and the last word of the pattern (an
 * aligned-block fingerprint only sees an end byte in one of its B alignments) and the word with the rarest byte of
 * the pattern (by the text's histogram).  Dense near-misses (a^(m-1)c in a^n: every position passes the
 * fingerprints) are rejected here 32 consecutive positions per load instruction instead of one warp-cooperative
 * verification each.  Returns 0 iff the words agree; an invalid candidate reads position 0 (always inside the
 * buffer) and reports a difference, so the loads of several candidates are issued back to back. */
struct ProbeWords {
    int k1, last;        // word index of the rarest byte (-1: none); index of the last word
    uint32_t x0, x1, xl; // pattern words: first, rarest, last
    uint32_t m1;         // mask of the rarest word (the last pattern word may be partial)
};

__device__ __forceinline__ ProbeWords probe_words(const ScanArgs &a, const uint32_t *pat)
{
    ProbeWords w;
    w.last = a.nwords - 1;
    w.k1 = a.probe < 0 ? -1 : (a.probe < w.last ? a.probe : w.last);
    w.x0 = __ldg(pat);
    w.x1 = __ldg(pat + (w.k1 < 0 ? 0 : w.k1));
    w.xl = __ldg(pat + w.last);
    w.m1 = w.k1 == w.last ? a.tailmask : 0xffffffffu;
    return w;
}

__device__ __forceinline__ uint32_t probe_diff(const ScanArgs &a, const ProbeWords &w, long long j, bool valid)
{
    const long long js = valid ? j : 0;
    uint32_t d = (text_window32(a.text, js + 4ll * w.last) ^ w.xl) & a.tailmask;
    if (w.k1 > 0 && w.k1 < w.last) d |= text_window32(a.text, js + 4ll * w.k1) ^ w.x1;
    if (w.last > 0) d |= text_window32(a.text, js) ^ w.x0;
    return valid ? d : 1u;
}

/* Linear-time handling of dense periodic matches (the job kmp.c:56-65 does with its failure function).
 * x has smallest period p with 2p <= m, and text[j .. j+m) == x.  While text[e] == text[e-p] holds from e = j+m on,
 * every j + i*p with j + i*p + m <= e is an occurrence, and NO other position inside the run is (a second
 * alignment would give x a period shorter than p).  Streams that comparison forward and returns how many further
 * occurrences j + i*p (i >= 1) start below own_end - the positions this warp's group owns.
 * From the first word boundary at or after j+m on, text[i..] is read as aligned words and text[i-p..] as a funnel
 * shift of two aligned words with ONE shift amount for the whole warp: 512 bytes per step, 12 coalesced loads a lane. */
__device__ __forceinline__ long long extend_run(const ScanArgs &a, long long j, long long own_end, unsigned lane)
{
    const long long p = a.period;
    if (own_end > a.hi) own_end = a.hi;
    if (j + p >= own_end) return 0;
    const long long lim = own_end - 1 + a.m;  // the text bytes that matter end here
    long long e = j + a.m;
    while ((e & 3) && e < lim && a.text[e] == a.text[e - p]) e++;  // up to 3 bytes (uniform loads) to a word boundary
    if ((e & 3) || e >= lim) {                                      // the run ends before it
        if (e > lim) e = lim;
        const long long k0 = (e - a.m - j) / p, kmax0 = (own_end - 1 - j) / p;
        return k0 < kmax0 ? k0 : kmax0;
    }
    const uint32_t *hi_w = reinterpret_cast<const uint32_t *>(a.text) + (e >> 2) + lane;       // text[e + 4*lane ..]
    const uint32_t *lo_w = reinterpret_cast<const uint32_t *>(a.text) + ((e - p) >> 2) + lane;  // text[e - p + 4*lane ..]
    const unsigned sh = (unsigned)((e - p) & 3) * 8u;
    while (e < lim) {
        int bad = 512;  // first differing byte among this lane's four words (word k of lane l: e + 128k + 4l)
#pragma unroll
        for (int k = 3; k >= 0; k--) {
            const uint32_t dd = __ldg(hi_w + 32 * k) ^ __funnelshift_r(__ldg(lo_w + 32 * k), __ldg(lo_w + 32 * k + 1), sh);
            if (dd) bad = 128 * k + 4 * (int)lane + ((__ffs(dd) - 1) >> 3);
        }
        const int first = __reduce_min_sync(kFull, bad);
        if (first == 512) {
            e += 512;
            hi_w += 128;
            lo_w += 128;
            continue;
        }
        e += first;
        break;
    }
    if (e > lim) e = lim;
    const long long k = (e - a.m - j) / p, kmax = (own_end - 1 - j) / p;
    return k < kmax ? k : kmax;
}

__device__ __forceinline__ void emit_match(const ScanArgs &a, long long j)
{
    if (a.pos) atomicOr(a.pos + (j >> 5), 1u << (j & 31));
}

__device__ __forceinline__ void block_reduce_count(const ScanArgs &a, unsigned cnt, unsigned nverified)
{
    __shared__ unsigned s_cnt[kThreads / 32];
    __shared__ unsigned s_ver[kThreads / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(kFull, cnt, o);
        nverified += __shfl_xor_sync(kFull, nverified, o);
    }
    if (lane == 0) {
        s_cnt[warp] = cnt;
        s_ver[warp] = nverified;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long c = 0, v = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) {
            c += s_cnt[w];
            v += s_ver[w];
        }
        if (c) atomicAdd(a.counters + 0, c);
        if (v) atomicAdd(a.counters + 2, v);
    }
}

/* ===================================================================== scan_block<B, L, U> */

struct CandResult {
    unsigned cnt, nver;
    long long run_j;  // start of the last verified occurrence seen by this warp (-1: none)
};
/* (run_j also marks how far the warp has DECIDED: candidates are visited in ascending order, and after a run of a
 * periodic pattern was counted arithmetically run_j is its last occurrence - every candidate <= run_j is settled) */

/* fingerprint of the L bytes starting at word q of W (L = 4, 8, 16 or 32; L == 32 only with q == 0) */
template <int L>
__device__ __forceinline__ uint32_t block_fp(const uint32_t *W, int q)
{
    if (L == 4) return W[q];
    if (L == 8) return fp8(W[q], W[q + 1]);
    if (L == 16) return fp16(W[q], W[q + 1], W[q + 2], W[q + 3]);
    return fp32(fp16(W[0], W[1], W[2], W[3]), fp16(W[4], W[5], W[6], W[7]));
}

/* B = 2 (windows at every EVEN text position, L = 3..8 bytes): fingerprint of the L bytes lo|hi, the bytes of
 * hi beyond L multiplied away (hi * 2^(8*(8-L)) keeps its low L-4 bytes): IMADs only, no mask */
template <int L>
__host__ __device__ __forceinline__ uint32_t pair_fp(uint32_t lo, uint32_t hi)
{
    if (L == 4) return lo;
    if (L == 3) return lo * 256u;  // three bytes: the fourth is multiplied away
    return lo * K3 + hi * (1u << (L > 4 ? 8 * (8 - L) : 0));
}

/* words a lane needs for its 32-byte vector: 8 own + halo from the next lane */
template <int B, int L>
struct BlockShape {
    static constexpr int HALO = B == 2 ? (L <= 6 ? 1 : 2) : (L - B) / 4;
    static constexpr int NWORDS = 8 + HALO;
};

/* fingerprint of block i (byte offset i*B) of a vector held in s[] */
template <int B, int L>
__device__ __forceinline__ uint32_t slot_fp(const uint32_t *s, int i)
{
    if (B == 2) {
        const int q = i >> 1;
        if (i & 1)
            return pair_fp<L>(__funnelshift_r(s[q], s[q + 1], 16), L > 4 ? __funnelshift_r(s[q + 1], s[q + 2], 16) : 0u);
        return pair_fp<L>(s[q], L > 4 ? s[q + 1] : 0u);
    }
    return block_fp<L>(s, i * (B / 4));
}

/* Slow path of scan_block: some lane of this warp saw a fingerprint hit in one of the 1 KiB chunks of its
 * group; cbase = byte offset of that chunk.  Nothing but scalars is passed in and the call is made after the
 * group's vectors are dead: every lane re-reads its 32-byte vector (+ halo) from memory - an L2 hit.  Steps:
 *   1. lane-local: which of the 32 (block, d) slots of the vector hit -> a candidate start position each;
 *   2. only when candidates are dense: three probe words per candidate, 32 consecutive candidates per load
 *      instruction - dense near-misses die here (skipped for periodic patterns, whose dense candidates are runs);
 *   3. warp-cooperative: the survivors, in ascending order, are verified exactly; after a verified occurrence of a
 *      periodic pattern the rest of its run is counted arithmetically (extend_run) and everything up to the last
 *      occurrence of the run is settled (run_j). */
/* Bit k set <=> the text word at byte jl + k + 4*widx equals x under `msk`, for the 32 consecutive start positions
 * jl .. jl+31 (any alignment, jl may lie before local byte 0: the zeroed front pad).  Ten aligned words are loaded,
 * realigned once by the dynamic byte offset, and the 32 windows come out of them at compile-time shifts. */
__device__ __forceinline__ unsigned probe_mask32(const uint8_t *text, long long jl, int widx, uint32_t x, uint32_t msk)
{
    const long long b = jl + 4ll * widx;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(text) + (b >> 2);  // (floor)
    const unsigned sh = (unsigned)(b & 3) * 8u;
    uint32_t T[10];
#pragma unroll
    for (int i = 0; i < 10; i++) T[i] = __ldg(p + i);
    uint32_t R[9];
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = __funnelshift_r(T[i], T[i + 1], sh);
    unsigned mk = 0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const uint32_t w = (k & 3) == 0 ? R[k >> 2] : __funnelshift_r(R[k >> 2], R[(k >> 2) + 1], 8 * (k & 3));
        mk |= (((w ^ x) & msk) == 0 ? 1u : 0u) << k;
    }
    return mk;
}

template <int B, int L, int GROUP>
__device__ __noinline__ CandResult block_candidates(const ScanArgs &a, long long cbase, unsigned lane, long long run_j)
{
    constexpr int NWORDS = BlockShape<B, L>::NWORDS;
    // GROUP = bytes of text one warp iteration owns: the candidates of this group start below own_end
    const long long own_end = (cbase / GROUP + 1) * GROUP - (B - 1) - a.anchor;
    unsigned cnt = 0, nver = 0;
    {
        const long long voff = cbase + lane * 32;
        // 1. own slots.  Slot sl = blk*B + dd stands for the candidate voff + blk*B - (B-1-dd) - anchor: ascending in sl
        unsigned own = 0;
        {
            uint32_t W[11] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            const uint32_t *w = reinterpret_cast<const uint32_t *>(a.text + voff);
#pragma unroll
            for (int i = 0; i < NWORDS; i++) W[i] = __ldg(w + i);
#pragma unroll
            for (int blk = 0; blk < 32 / B; blk++) {
                const uint32_t h = slot_fp<B, L>(W, blk);
#pragma unroll
                for (int dd = 0; dd < B; dd++) own |= (h == a.H[B - 1 - dd] ? 1u : 0u) << (blk * B + dd);
            }
        }
        // 2. probe words - only when candidates are DENSE.  A handful of candidates (the normal case) and the
        //    candidates of a periodic pattern (runs) go straight to step 3 ("direct").  Otherwise every lane tests
        //    one pattern word against ALL 32 of its slots at once: the slots of a lane are 32 consecutive start
        //    positions, so their words are the 32 byte-shifted windows of ten text words (probe_mask32).  The most
        //    selective word first - the one with the text's rarest byte, else the last one (the fingerprints sit at
        //    the other end for prefix anchors) - and the next word only while some candidate of the chunk is
        //    still alive (a warp vote): a text that looks like the pattern everywhere except in one place costs
        //    one probe per chunk, not three per candidate.
        const bool direct = a.period != 0 || __reduce_add_sync(kFull, (unsigned)__popc(own)) <= 4u;
        if (!direct) {
            const ProbeWords pw = probe_words(a, a.pat);
            const long long jl = voff - (B - 1) - a.anchor;  // slot sl of this lane is the start position jl + sl
            const bool use1 = pw.k1 > 0 && pw.k1 < pw.last, use0 = pw.last > 0;
            if (use1) own &= probe_mask32(a.text, jl, pw.k1, pw.x1, 0xffffffffu);
            if (!use1 || __any_sync(kFull, own != 0)) own &= probe_mask32(a.text, jl, pw.last, pw.xl, a.tailmask);
            if (use0 && __any_sync(kFull, own != 0)) own &= probe_mask32(a.text, jl, 0, pw.x0, 0xffffffffu);
        }
        // 3. exact verification of the survivors, vector by vector, slot by slot: ascending start positions
        unsigned pending = __ballot_sync(kFull, own != 0);  // bit r: vector r
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const long long soff = cbase + src * 32;
            if (cbase + 1024 - B - a.anchor <= run_j) break;  // a counted run settled the rest of the chunk
            if (soff + 32 - B - a.anchor <= run_j) continue;
            unsigned slots = __shfl_sync(kFull, own, src);
            while (slots) {
                const int sl = __ffs(slots) - 1;
                slots &= slots - 1;
                const long long j = soff + (sl / B) * B - (B - 1 - sl % B) - a.anchor;
                if (j <= run_j || j < 0 || j >= a.hi) continue;  // settled by a counted run, or not an owned start
                nver += (lane == 0);
                // Period-aware verification: x has period p and text[run_j .. run_j+m) == x.  If j = run_j + k*p
                // lies inside that occurrence, text[j .. run_j+m) equals x[0 .. m-k*p) by periodicity and only
                // the last k*p bytes remain to be compared (rounded down to a word).
                int from_word = 0;
                const long long delta = j - run_j;
                if (a.period && run_j >= 0 && delta < a.m && delta % a.period == 0) from_word = (a.m - (int)delta) >> 2;
                if (warp_verify(a, j, lane, nullptr, from_word)) {
                    run_j = j;
                    if (lane == 0) {
                        cnt++;
                        emit_match(a, j);
                    }
                    if (a.period) {  // count the rest of a periodic run arithmetically
                        const long long k = extend_run(a, j, own_end, lane);
                        if (k > 0) {
                            if (lane == 0) cnt += (unsigned)k;
                            if (a.pos)
                                for (long long i = 1 + lane; i <= k; i += 32) emit_match(a, j + i * a.period);
                            run_j = j + k * a.period;
                        }
                    }
                }
            }
        }
    }
    CandResult r;
    r.cnt = cnt;
    r.nver = nver;
    r.run_j = run_j;
    return r;
}

/* B = alignment of the fingerprinted blocks = number of pattern alignments compared per block;
 * L >= B = bytes fingerprinted from each block start.  L == B: disjoint blocks (no halo).  L > B: overlapping
 * windows — the selectivity of L bytes for patterns as short as L+B-1 (e.g. B=4, L=8 from m = 11). */
template <int B, int L>
__device__ __forceinline__ bool block_hit(const ScanArgs &a, const uint32_t *W)
{
    bool hit = false;
    if (B == 2) {  // windows at even offsets: the odd half-words by one funnel shift per word
        uint32_t F[9];
#pragma unroll
        for (int k = 0; k < 9; k++) F[k] = (L > 4 || k < 8) ? __funnelshift_r(W[k], W[k + 1], 16) : 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t h0 = pair_fp<L>(W[k], W[k + 1]);
            const uint32_t h2 = pair_fp<L>(F[k], F[k + 1]);
            hit |= (h0 == a.H[0]) | (h0 == a.H[1]) | (h2 == a.H[0]) | (h2 == a.H[1]);
        }
    } else if (L == 16 && B == 4) {  // share the 8-byte halves between the overlapping 16-byte windows
        uint32_t g8[10];
#pragma unroll
        for (int q = 0; q < 10; q++) g8[q] = fp8(W[q], W[q + 1]);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t h = g8[q] * K2 + g8[q + 2];
#pragma unroll
            for (int d = 0; d < 4; d++) hit |= (h == a.H[d]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 32 / B; k++) {
            const uint32_t h = block_fp<L>(W, k * (B / 4));
#pragma unroll
            for (int d = 0; d < B; d++) hit |= (h == a.H[d]);
        }
    }
    return hit;
}

/* The hot loop only NOTES that a chunk needs the slow path; the chunks are resolved after the group's last chunk,
 * when no loaded vector is live any more (the slow path needs many registers: called mid-group it made the hot loop
 * spill).  Deliberately a real call: the rarely taken branch around it ends the basic block of each chunk, which
 * keeps the chunks' instruction streams in load order (one vote per group instead measured 2-19% slower). */
__device__ __noinline__ unsigned note_hit(unsigned pend, int c) { return pend | (1u << c); }

/* group = U x 1 KiB per warp */
template <int B, int L, int U>
__global__ void __launch_bounds__(kThreads, 3) scan_block_kernel(const __grid_constant__ ScanArgs a)
{
    constexpr int HALO = BlockShape<B, L>::HALO;  // words needed from the next lane
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * kThreads) >> 5;
    unsigned cnt = 0, nver = 0;
    long long run_j = -1;  // last verified occurrence of this warp, for period-aware verification
    unsigned pend = 0;     // bit c: chunk c of the current group passed the filter in some lane (warp-uniform)

    for (unsigned long long g = warp; g < a.ngroups; g += nwarps) {
        const uint8_t *p = a.text + g * (U * 1024ull) + lane * 32;
        V8 v[U];
#pragma unroll
        for (int c = 0; c < U; c++) v[c] = ldg_stream256(p + c * 1024);
        uint32_t gnw[4] = {0u, 0u, 0u, 0u};
        if (HALO > 0 && lane == 0) {  // halo of the last chunk's lane 31: first words of the next group
            const uint4 gn = __ldg(reinterpret_cast<const uint4 *>(a.text + (g + 1) * (U * 1024ull)));
            gnw[0] = gn.x;
            gnw[1] = gn.y;
            gnw[2] = gn.z;
            gnw[3] = gn.w;
        }
#pragma unroll
        for (int c = 0; c < U; c++) {
            uint32_t W[11];
#pragma unroll
            for (int i = 0; i < 8; i++) W[i] = v[c].w[i];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                W[8 + i] = 0u;
                if (i < HALO) {
                    const uint32_t nxt = (c + 1 < U) ? v[(c + 1 < U) ? c + 1 : c].w[i] : gnw[i];
                    W[8 + i] = __shfl_sync(kFull, lane == 0 ? nxt : v[c].w[i], (lane + 1) & 31);
                }
            }
            if (__any_sync(kFull, block_hit<B, L>(a, W))) pend = note_hit(pend, c);
        }
        if (pend) {  // rare: resolve the noted chunks in ascending order
#pragma unroll 1
            for (int c = 0; c < U; c++) {
                const long long cbase = (long long)(g * (U * 1024ull)) + c * 1024;
                if (((pend >> c) & 1u) && cbase + 1024 - B - a.anchor > run_j) {  // else: settled by a counted run
                    const CandResult r = block_candidates<B, L, U * 1024>(a, cbase, lane, run_j);
                    cnt += r.cnt;
                    nver += r.nver;
                    run_j = r.run_j;
                }
            }
            pend = 0;
        }
    }
    block_reduce_count(a, cnt, nver);
}


/* ===================================================================== scan_block_tma<B> (experiment) */

/* The same aligned-block filter fed by TMA bulk copies instead of per-thread loads: one elected producer thread
 * streams 16 KiB tiles global -> shared with cp.async.bulk (UBLKCP) into a ring of kTmaStages buffers,
 * completion signalled on an mbarrier per stage; 8 consumer warps read their 2 KiB slice of the tile with
 * LDS.128, run block_hit / block_candidates, and release the stage through a second mbarrier.
 * Kept to MEASURE the claim of DESIGN.md section 3.3 (no reuse => shared-memory staging cannot beat register
 * streaming); selected only with SMART_B200_TMA=1. */
constexpr int kTmaTile = 16384;
constexpr int kTmaStages = 2;  // default ring depth x 6 CTAs/SM (a.stride overrides): best of the r1 sweep
constexpr int kTmaThreads = 288;  // 8 consumer warps + 1 producer warp

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_load(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int B>
__global__ void __launch_bounds__(kTmaThreads, 1) scan_block_tma_kernel(const __grid_constant__ ScanArgs a)
{
    extern __shared__ __align__(128) uint8_t tma_smem[];
    uint8_t *tiles = tma_smem;
    const int stages = a.stride;                 // ring depth
    const unsigned long long limit = a.nsamples;  // readable bytes of the shard buffer (multiple of 16)
    uint64_t *full = reinterpret_cast<uint64_t *>(tma_smem + stages * kTmaTile);
    uint64_t *empty = full + stages;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long ntiles = a.ngroups;  // here: number of 16 KiB tiles
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned cnt = 0, nver = 0;
    long long run_j = -1;
    if (warp == 8) {
        if (lane == 0) {  // producer
            unsigned k = 0;
            for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++) {
                const int s = k % stages;
                const unsigned round = k / stages;
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                const unsigned long long off = tile * kTmaTile;
                const unsigned bytes = (unsigned)min((unsigned long long)kTmaTile, limit - off);
                mbar_expect_tx(&full[s], bytes);  // a short last tile leaves stale bytes behind it: they can only
                tma_bulk_load(tiles + s * kTmaTile, a.text + off, bytes, &full[s]);  // add candidates >= hi
            }
        }
    } else {
        unsigned k = 0;
        for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++) {
            const int s = k % stages;
            mbar_wait(&full[s], (k / stages) & 1);
            const uint8_t *slice = tiles + s * kTmaTile + warp * 2048;
            V8 v[2];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const uint4 lo = *reinterpret_cast<const uint4 *>(slice + c * 1024 + lane * 32);
                const uint4 hi = *reinterpret_cast<const uint4 *>(slice + c * 1024 + lane * 32 + 16);
                v[c].w[0] = lo.x; v[c].w[1] = lo.y; v[c].w[2] = lo.z; v[c].w[3] = lo.w;
                v[c].w[4] = hi.x; v[c].w[5] = hi.y; v[c].w[6] = hi.z; v[c].w[7] = hi.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);  // the slice is in registers: the stage may be refilled
#pragma unroll
            for (int c = 0; c < 2; c++) {
                if (__any_sync(kFull, block_hit<B, B>(a, v[c].w))) {
                    const long long cbase = (long long)(tile * kTmaTile) + warp * 2048 + c * 1024;  // 2 KiB slice per warp
                    const CandResult r = block_candidates<B, B, 2048>(a, cbase, lane, run_j);
                    cnt += r.cnt;
                    nver += r.nver;
                    run_j = r.run_j;
                }
            }
        }
    }
    // block_reduce_count expects kThreads; reduce here over the 9 warps
    __shared__ unsigned t_cnt[9], t_ver[9];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(kFull, cnt, o);
        nver += __shfl_xor_sync(kFull, nver, o);
    }
    if (lane == 0) {
        t_cnt[warp] = cnt;
        t_ver[warp] = nver;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long c = 0, v = 0;
        for (int w = 0; w < 9; w++) {
            c += t_cnt[w];
            v += t_ver[w];
        }
        if (c) atomicAdd(a.counters + 0, c);
        if (v) atomicAdd(a.counters + 2, v);
    }
}

/* ===================================================================== scan_window<NW, OFF, DENSE, U> */

/* Per-lane slow path of scan_window: `mm` has bit p set for every start position p of this lane's
 * 32-byte vector whose window words match exactly; verify the pattern words from `from` on and count. */
__device__ __noinline__ uint2 window_matches(const ScanArgs &a, unsigned mm, long long voff, int from)
{
    unsigned cnt = 0, nver = 0;
    while (mm) {
        const int p = __ffs(mm) - 1;
        mm &= mm - 1;
        const long long j = voff + p;
        if (j >= a.hi) break;
        if (a.nwords > from) {
            nver++;
            // the last word first: a near-miss at the far end must not cost a walk over the whole pattern
            const int last = a.nwords - 1;
            if ((text_window32(a.text, j + 4ll * last) ^ __ldg(a.pat + last)) & a.tailmask) continue;
            if (!lane_verify(a, j, from)) continue;
        }
        cnt++;
        emit_match(a, j);
    }
    return make_uint2(cnt, nver);
}

/* Window fingerprint at EVERY start position p: NW 32-bit words at byte offsets p, p+4, .., the last one at
 * p + 4*(NW-2) + OFF.  OFF = 4: an 8/12/16-byte prefix window (the rest of a longer pattern is verified on
 * a hit).  OFF < 4: m = 4*(NW-1) + OFF, the last word overlaps the previous one and the window covers the
 * pattern EXACTLY, so a window match is an occurrence.  Wider windows are for low-entropy texts (sigma=2)
 * where 8 bytes are not selective.  DENSE (window == pattern, frequent matches): every position is compared
 * exactly and counted in registers, no branch. */
/* register caps (measured, sigma=4 m=8 and sigma=2 m=16 on 4 GiB): unconstrained the kernels take 83+
 * registers = 2 CTAs/SM (5.7 / 5.1 TB/s); two-word windows fit 64 registers = 4 CTAs/SM (6.7 TB/s), the
 * 12/16-byte windows spill at 64 and are best at 80 registers = 3 CTAs/SM (5.45 TB/s) */
template <int NW, int OFF, bool DENSE, int U>
__global__ void __launch_bounds__(kThreads, NW == 2 ? 4 : 3) scan_window_kernel(const __grid_constant__ ScanArgs a)
{
    constexpr int LAST = 4 * (NW - 2) + OFF;     // byte offset of the last window word
    constexpr int NS = 32 + LAST;                // shifted words needed: byte offsets 0 .. 31+LAST
    constexpr int NWORDS = (NS + 2) / 4 + 1;     // 32-bit words feeding them (8 own + halo)
    constexpr int HALO = NWORDS - 8;
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * kThreads) >> 5;
    unsigned cnt = 0, nver = 0;

    for (unsigned long long g = warp; g < a.ngroups; g += nwarps) {
        const uint8_t *p = a.text + g * (U * 1024ull) + lane * 32;
        V8 v[U];
#pragma unroll
        for (int c = 0; c < U; c++) v[c] = ldg_stream256(p + c * 1024);
        // halo of the last chunk's lane 31: first 16 bytes of the next group, fetched by lane 0 (L2 hit later)
        uint4 gn = make_uint4(0u, 0u, 0u, 0u);
        if (lane == 0) gn = __ldg(reinterpret_cast<const uint4 *>(a.text + (g + 1) * (U * 1024ull)));
        const uint32_t gnw[4] = {gn.x, gn.y, gn.z, gn.w};
#pragma unroll
        for (int c = 0; c < U; c++) {
            // the bytes after this lane's vector: lane+1's first words; lane 31 takes lane 0 of the next chunk
            uint32_t W[NWORDS];
#pragma unroll
            for (int i = 0; i < 8; i++) W[i] = v[c].w[i];
#pragma unroll
            for (int i = 0; i < HALO; i++) {
                const uint32_t nxt = (c + 1 < U) ? v[(c + 1 < U) ? c + 1 : c].w[i] : gnw[i];
                const uint32_t snd = lane == 0 ? nxt : v[c].w[i];
                W[8 + i] = __shfl_sync(kFull, snd, (lane + 1) & 31);
            }
            // shifted words for byte offsets 0 .. 31+LAST.  (Forming them on the idle FMA pipe instead, as
            // hi*2^k + umulhi(lo, 2^k), was measured slower: IMAD.HI is not full rate.)
            uint32_t S[NS];
#pragma unroll
            for (int b = 0; b < NS; b++)
                S[b] = (b & 3) ? __funnelshift_r(W[b >> 2], W[(b >> 2) + 1], 8 * (b & 3)) : W[b >> 2];
            const long long voff = (long long)(g * (U * 1024ull)) + c * 1024 + lane * 32;
            if (DENSE && voff + 32 <= a.hi) {
#pragma unroll
                for (int b = 0; b < 32; b++) {
                    bool eq = (S[b] == a.X0) && (S[b + LAST] == a.Xw[NW - 2]);
                    if (NW >= 3) eq = eq && (S[b + 4] == a.Xw[0]);
                    if (NW >= 4) eq = eq && (S[b + 8] == a.Xw[1]);
                    cnt += eq;
                }
                continue;
            }
            bool hit = DENSE;
            if (!DENSE) {
                if (NW == 2) {
#pragma unroll
                    for (int b = 0; b < 32; b++) hit |= (fp8(S[b], S[b + LAST]) == a.G);
                } else if (NW == 3) {
#pragma unroll
                    for (int b = 0; b < 32; b++) hit |= (fp8(S[b], S[b + 4]) * K2 + S[b + LAST] == a.G);
                } else if (OFF == 4) {
                    uint32_t g8[40];
#pragma unroll
                    for (int b = 0; b < 40; b++) g8[b] = fp8(S[b], S[b + 4]);
#pragma unroll
                    for (int b = 0; b < 32; b++) hit |= (g8[b] * K2 + g8[b + 8] == a.G);
                } else {
#pragma unroll
                    for (int b = 0; b < 32; b++)
                        hit |= (fp8(S[b], S[b + 4]) * K2 + fp8(S[b + 8], S[b + LAST]) == a.G);
                }
            }
            if (hit && voff < a.hi) {  // per lane: which positions match all window words exactly
                unsigned mm = 0;
#pragma unroll
                for (int b = 0; b < 32; b++) {
                    bool eq = (S[b] == a.X0) && (S[b + LAST] == a.Xw[NW - 2]);
                    if (NW >= 3) eq = eq && (S[b + 4] == a.Xw[0]);
                    if (NW >= 4) eq = eq && (S[b + 8] == a.Xw[1]);
                    mm |= (eq ? 1u : 0u) << b;
                }
                if (mm) {
                    const uint2 r = window_matches(a, mm, voff, NW);
                    cnt += r.x;
                    nver += r.y;
                }
            }
        }
    }
    block_reduce_count(a, cnt, nver);
}


/* ===================================================================== scan_skip<G> */

/* Sub-linear scan for the skip families (BM, Horspool, QS, Raita, Tuned-BM, HASHq, QF are sub-linear on the
 * CPU for the same reason): every occurrence [j, j+m) contains the G-byte gram at the first multiple t of
 * `stride` with t >= j + anchor, provided stride + G - 1 <= m - anchor.  So only ONE 32-byte sector out of
 * every stride/32 is read.  The gram's fingerprint is looked up in a Bloom bitmap of the `stride` pattern
 * grams it could be (shared memory); a hit is resolved against the fingerprint list and verified exactly.
 * Text GB/s (n / t) can therefore exceed the read-once HBM roofline.  Measured (ncu, profiles/): DRAM
 * delivers 128 bytes per sampled sector, so the traffic is ~n*128/stride and skipping pays from stride 192. */
#ifndef SB_SKIP_L2
#define SB_SKIP_L2 ".L2::64B" /* smallest L2 prefetch-size hint PTX offers: the sampled sectors are isolated */
#endif
template <int G>
__device__ __forceinline__ uint32_t skip_fp(const uint8_t *p)
{
    if (G == 8) {
        uint2 v;
        asm volatile("ld.global.nc.L1::no_allocate" SB_SKIP_L2 ".v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
        return fp8(v.x, v.y);
    } else if (G == 16) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate" SB_SKIP_L2 ".v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(p));
        return fp16(v.x, v.y, v.z, v.w);
    } else {
        V8 v;
        asm volatile("ld.global.nc.L1::no_allocate" SB_SKIP_L2 ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]),
                       "=r"(v.w[7])
                     : "l"(p));
        return fp32(fp16(v.w[0], v.w[1], v.w[2], v.w[3]), fp16(v.w[4], v.w[5], v.w[6], v.w[7]));
    }
}

__device__ __noinline__ uint2 skip_candidates(const ScanArgs &a, uint32_t h, long long t, bool hit, unsigned lane)
{
    unsigned cnt = 0, nver = 0;
    const uint2 *table = reinterpret_cast<const uint2 *>(a.aux + kBloomWords);
    const ProbeWords pw = probe_words(a, a.pat);
    unsigned pending = __ballot_sync(kFull, hit);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const uint32_t sh = __shfl_sync(kFull, h, src);
        const long long st = __shfl_sync(kFull, t, src);
        // 32 consecutive table entries per round; the probe chain ends at the first empty entry
        const uint32_t base = table_slot(sh);
        for (int round = 0; round < kSkipTableSize / 32; round++) {
            const uint2 e = __ldg(table + ((base + round * 32 + lane) & (kSkipTableSize - 1)));
            const unsigned empty = __ballot_sync(kFull, e.y == 0);
            // every matching entry is an alignment d: the lane that holds it pre-checks its candidate (probe words)
            const long long jl = st - ((int)e.y - 1) - a.anchor;
            bool cand = e.y != 0 && e.x == sh && jl >= 0 && jl < a.hi;
            cand = probe_diff(a, pw, jl, cand) == 0;
            unsigned eq = __ballot_sync(kFull, cand);
            if (empty) eq &= (1u << (__ffs(empty) - 1)) - 1u;
            while (eq) {
                const int l = __ffs(eq) - 1;
                eq &= eq - 1;
                const long long j = __shfl_sync(kFull, jl, l);
                nver += (lane == 0);
                if (warp_verify(a, j, lane) && lane == 0) {
                    cnt++;
                    emit_match(a, j);
                }
            }
            if (empty) break;
        }
    }
    return make_uint2(cnt, nver);
}

template <int G>
__global__ void __launch_bounds__(kThreads) scan_skip_kernel(const __grid_constant__ ScanArgs a)
{
    constexpr int U = 8;  // independent sector loads in flight per thread
    __shared__ uint32_t bloom[kBloomWords];
    for (int i = threadIdx.x; i < kBloomWords; i += kThreads) bloom[i] = __ldg(a.aux + i);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * kThreads) >> 5;
    const unsigned long long stride = (unsigned long long)a.stride;
    unsigned cnt = 0, nver = 0;

    for (unsigned long long k0 = warp * (32 * U); k0 < a.nsamples; k0 += nwarps * (32 * U)) {
        uint32_t h[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned long long k = k0 + u * 32 + lane;  // samples past the end read zero padding
            h[u] = skip_fp<G>(a.text + (k < a.nsamples ? k : a.nsamples - 1) * stride);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned long long k = k0 + u * 32 + lane;
            const uint32_t s1 = bloom_slot(h[u]), s2 = bloom_slot2(h[u]);
            const bool hit = ((bloom[s1 >> 5] >> (s1 & 31)) & (bloom[s2 >> 5] >> (s2 & 31)) & 1u) && k < a.nsamples;
            if (__any_sync(kFull, hit)) {
                const uint2 r = skip_candidates(a, h[u], (long long)(k * stride), hit, lane);
                cnt += r.x;
                nver += r.y;
            }
        }
    }
    block_reduce_count(a, cnt, nver);
}


/* ===================================================================== scan_packed<K> (alphabets of <= 4 symbols) */

/* A text over at most four distinct byte values (DNA, sigma <= 4 generators) is also kept as a 2-bit recoding
 * (1 bit for two symbols): character i lives in word i/CPW at bits BITS*(i%CPW).., CPW = 32/BITS (pack_kernel,
 * once per text).  A word is then an EXACT CPW-character fingerprint, and the scan reads 1/4 (1/8) of the bytes.
 * Stated for CPW = 16:
 *   every occurrence of length m >= 16K+15 contains the 16 characters of some word whose index is a multiple of K;
 *   that word equals one of the 16K words packed from x[anchor+d .. anchor+d+16), d < 16K.
 * So the kernel streams the packed buffer (256-bit loads, as scan_block), looks every K-th word up in the Bloom
 * bitmap of those 16K words (shared memory, 1 probe) and sends the rare hits through the exact table and the
 * byte-wise verification of scan_skip (skip_candidates: positions and counts come from the ORIGINAL text).
 * Text rate = 4 x the packed-byte rate: above the read-once HBM roofline of the byte text, like scan_skip. */
struct PackLut {
    uint32_t sym[16];  // byte values of codes 0..15 (an unused code holds a value no byte has)
};

template <int BITS>
__device__ __forceinline__ uint32_t pack_code(uint32_t b, const PackLut &lut, bool &ok)
{
    uint32_t c = 0;
#pragma unroll
    for (int i = 1; i < (1 << BITS); i++)
        if (b == lut.sym[i]) c = i;
    ok = (b == lut.sym[c]);
    return c;
}

/* Packs the text at BITS per character: word w holds the CPW = 32/BITS characters text[CPW*w .. CPW*(w+1)),
 * character i at bits BITS*i...  Each thread turns 32 text bytes into 32/CPW words.  Bytes at or beyond `len`
 * pack as code 0; *bad is set if a byte below `len` is none of the symbols (the caller then drops the packed
 * copy).  BITS = 1 for two-symbol texts, 2 for up to four symbols, 4 for up to sixteen. */
template <int BITS>
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *text, unsigned long long len, unsigned long long cap,
                                                   unsigned long long nwords, uint32_t *out, const PackLut lut,
                                                   unsigned *bad)
{
    constexpr int CPW = 32 / BITS, WPU = 32 / CPW;  // characters per word, words per 32-byte unit
    bool any_bad = false;
    const unsigned long long nunits = nwords / WPU;
    for (unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < nunits;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t word[WPU];
#pragma unroll
        for (int k = 0; k < WPU; k++) word[k] = 0;
        if (32 * u + 32 <= cap) {
#pragma unroll
            for (int v16 = 0; v16 < 2; v16++) {
                const uint4 v = ldg_stream128(text + 32 * u + 16 * v16);
                const uint32_t src[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    bool ok;
                    const uint32_t c = pack_code<BITS>((src[i >> 2] >> (8 * (i & 3))) & 0xffu, lut, ok);
                    const int ch = 16 * v16 + i;  // character index inside the unit
                    if (32 * u + ch < len) {
                        word[ch / CPW] |= c << (BITS * (ch % CPW));
                        any_bad |= !ok;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < WPU; k++) out[u * WPU + k] = word[k];
    }
    if (any_bad) atomicOr(bad, 1u);
}

template <int K>
__global__ void __launch_bounds__(kThreads) scan_packed_kernel(const __grid_constant__ ScanArgs a)
{
    constexpr int U = 4;
    __shared__ uint32_t bloom[kBloomWords];
    for (int i = threadIdx.x; i < kBloomWords; i += kThreads) bloom[i] = __ldg(a.aux + i);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * kThreads) >> 5;
    const uint8_t *pk = reinterpret_cast<const uint8_t *>(a.packed);
    unsigned cnt = 0, nver = 0;

    for (unsigned long long g = warp; g < a.ngroups; g += nwarps) {
        const uint8_t *p = pk + g * (U * 1024ull) + lane * 32;
        V8 v[U];
#pragma unroll
        for (int c = 0; c < U; c++) v[c] = ldg_stream256(p + c * 1024);
#pragma unroll
        for (int c = 0; c < U; c++) {
            const unsigned long long q0 = (g * (U * 1024ull) + c * 1024 + lane * 32) >> 2;  // word index of w[0]
#pragma unroll
            for (int k = 0; k < 8; k += K) {  // q0 is a multiple of 8: word q0+k has index = k (mod K)
                // ONE probe: at most 128 words in 2^18 bits give < 5e-4 false hits per word, and the exact table
                // behind skip_candidates settles them; the second probe cost more than it saved.  (Also measured: a
                // bank-conflict-free 4096-bit pre-filter in 32 copies in front of this probe - slower, 14.5 -> 11.6
                // T characters/s at K = 1: the variant is issue-bound, not bank-conflict-bound.)
                const uint32_t w = v[c].w[k];
                const uint32_t s1 = bloom_slot(w);
                const bool hit = ((bloom[s1 >> 5] >> (s1 & 31)) & 1u) && q0 + k < a.nsamples;
                if (__any_sync(kFull, hit)) {
                    const uint2 r = skip_candidates(a, w, (long long)((q0 + k) * (unsigned long long)a.cpw), hit, lane);
                    cnt += r.x;
                    nver += r.y;
                }
            }
        }
    }
    block_reduce_count(a, cnt, nver);
}

/* Shorter patterns on the recoding (CPW-1 <= m < 2*CPW-1): a 16-bit half-word is an exact CPW/2-character
 * fingerprint and indexes a 65536-bit bitmap DIRECTLY (8 KiB of shared memory, no hashing).  KH = 1: both halves
 * of every word (m >= 15, 8 pattern alignments); KH = 2: the low half only (m >= 23, 16 alignments). */
constexpr int kHalfMapWords = 65536 / 32;

template <int KH>
__global__ void __launch_bounds__(kThreads) scan_packed_half_kernel(const __grid_constant__ ScanArgs a)
{
    constexpr int U = 4;
    __shared__ uint32_t bitmap[kHalfMapWords];
    for (int i = threadIdx.x; i < kHalfMapWords; i += kThreads) bitmap[i] = __ldg(a.aux + i);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * kThreads) >> 5;
    const uint8_t *pk = reinterpret_cast<const uint8_t *>(a.packed);
    unsigned cnt = 0, nver = 0;

    for (unsigned long long g = warp; g < a.ngroups; g += nwarps) {
        const uint8_t *p = pk + g * (U * 1024ull) + lane * 32;
        V8 v[U];
#pragma unroll
        for (int c = 0; c < U; c++) v[c] = ldg_stream256(p + c * 1024);
#pragma unroll
        for (int c = 0; c < U; c++) {
            const unsigned long long q0 = (g * (U * 1024ull) + c * 1024 + lane * 32) >> 2;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t w = v[c].w[k];
#pragma unroll
                for (int half = 0; half < (KH == 1 ? 2 : 1); half++) {
                    const uint32_t h = half ? (w >> 16) : (w & 0xffffu);
                    const bool hit = ((bitmap[h >> 5] >> (h & 31)) & 1u) && q0 + k < a.nsamples;
                    if (__any_sync(kFull, hit)) {
                        const uint2 r = skip_candidates(
                            a, h, (long long)((q0 + k) * (unsigned long long)a.cpw + half * (a.cpw >> 1)), hit, lane);
                        cnt += r.x;
                        nver += r.y;
                    }
                }
            }
        }
    }
    block_reduce_count(a, cnt, nver);
}

/* ===================================================================== scan_pswar<BITS, KF, EXACT> */

/* Short patterns on the 1-/2-bit recoding, SIMD within a register: CPW = 32/BITS start positions per instruction.
 * For packed word q (characters CPW*q ..) and pattern character k, the word S_k = funnelshift(W[q], W[q+1], BITS*k)
 * holds text character p+k at the lane (bit field) of start position p, so
 *     BITS = 1:  acc  &= ~(S_k ^ splat(x[k]))     one LOP3 per pattern character, acc = match mask
 *     BITS = 2:  diff |=  (S_k ^ splat(x[k]))     one LOP3; a lane of diff is 00 iff all characters matched
 * cost 1 SHF + 1 LOP3 per pattern character for 32 (sigma = 2) or 16 (sigma <= 4) start positions, on 1/8 or 1/4 of
 * the bytes: the per-position window kernels this replaces spent a SHF + IMAD + ISETP per POSITION (ALU-bound at
 * 3.5-5.2 TB/s).  KF = number of pattern characters in the chain (KF*BITS <= 32, one halo word).
 *   EXACT (KF == m): the mask IS the set of occurrences: counted with POPC in registers, nothing to verify
 *                    (this is also the linear-time path for texts where every position matches);
 *   !EXACT (KF < m): the chain is a KF-character filter (chosen to carry >= 20 bits) on x[anchor .. anchor+KF);
 *                    surviving positions are verified on the byte text.
 * a.H[k] = splat of the code of x[anchor+k]; a.stride = KF (unused by the kernel, kept for the launch arithmetic). */
template <int BITS>
__device__ __noinline__ uint2 pswar_matches(const ScanArgs &a, uint32_t mask, long long pbase, bool exact)
{
    unsigned cnt = 0, nver = 0;
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const long long j = pbase + b / BITS - a.anchor;
        if (j < 0 || j >= a.hi) continue;
        if (!exact) {
            nver++;
            if (!lane_verify(a, j, 0)) continue;
        }
        cnt++;
        emit_match(a, j);
    }
    return make_uint2(cnt, nver);
}

template <int BITS, int KF, bool EXACT>
__global__ void __launch_bounds__(kThreads) scan_pswar_kernel(const __grid_constant__ ScanArgs a)
{
    static_assert(BITS * KF <= 32 && KF >= 1, "the chain must fit one halo word");
    constexpr int U = 4;
    constexpr int CPW = 32 / BITS;
    constexpr uint32_t LANES = BITS == 1 ? 0xffffffffu : 0x55555555u;  // bit 0 of every character lane
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * kThreads) >> 5;
    const uint8_t *pk = reinterpret_cast<const uint8_t *>(a.packed);
    unsigned cnt = 0, nver = 0;

    for (unsigned long long g = warp; g < a.ngroups; g += nwarps) {
        const uint8_t *p = pk + g * (U * 1024ull) + lane * 32;
        V8 v[U];
#pragma unroll
        for (int c = 0; c < U; c++) v[c] = ldg_stream256(p + c * 1024);
        uint32_t gn = 0;
        if (KF > 1 && lane == 0) gn = __ldg(reinterpret_cast<const uint32_t *>(pk + (g + 1) * (U * 1024ull)));
#pragma unroll
        for (int c = 0; c < U; c++) {
            uint32_t W[9];
#pragma unroll
            for (int i = 0; i < 8; i++) W[i] = v[c].w[i];
            W[8] = 0u;
            if (KF > 1) {
                const uint32_t nxt = (c + 1 < U) ? v[(c + 1 < U) ? c + 1 : c].w[0] : gn;
                W[8] = __shfl_sync(kFull, lane == 0 ? nxt : v[c].w[0], (lane + 1) & 31);
            }
            // first text character of this lane's vector, and whether all its start positions are owned
            const long long pbase = (long long)(((g * (U * 1024ull) + c * 1024 + lane * 32) >> 2) * CPW);
            const bool inside = pbase - a.anchor >= 0 && pbase + 8 * CPW - a.anchor <= a.hi;
            uint32_t mk[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (BITS == 1) {
                    uint32_t acc = ~(W[q] ^ a.H[0]);
#pragma unroll
                    for (int k = 1; k < KF; k++) acc &= ~(__funnelshift_r(W[q], W[q + 1], k) ^ a.H[k]);
                    mk[q] = acc;
                } else {
                    uint32_t diff = W[q] ^ a.H[0];
#pragma unroll
                    for (int k = 1; k < KF; k++) diff |= __funnelshift_r(W[q], W[q + 1], BITS * k) ^ a.H[k];
                    mk[q] = ~(diff | (diff >> 1)) & LANES;
                }
            }
            if (EXACT && inside && !a.pos) {
#pragma unroll
                for (int q = 0; q < 8; q++) cnt += __popc(mk[q]);
                continue;
            }
            uint32_t any = mk[0];
#pragma unroll
            for (int q = 1; q < 8; q++) any |= mk[q];
            if (any != 0u && pbase - a.anchor < a.hi) {
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (mk[q]) {
                        const uint2 r = pswar_matches<BITS>(a, mk[q], pbase + q * CPW, EXACT);
                        cnt += r.x;
                        nver += r.y;
                    }
                }
            }
        }
    }
    block_reduce_count(a, cnt, nver);
}

/* ===================================================================== scan_multi<B> */

/* Multi-pattern scan: ALL `npat` patterns of one length are counted in ONE pass over
 * the text.  Same aligned-block argument as scan_block, but instead of comparing a block fingerprint with
 * the B fingerprints of one pattern it is looked up in a 3-probe Bloom bitmap (shared memory) holding the
 * npat*B fingerprints of all patterns; hits are resolved in the exact table {fingerprint, (pattern, d)} and
 * verified against that pattern; counts are kept per pattern. */
template <int B>
__device__ __noinline__ unsigned multi_candidates(const ScanArgs &a, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3,
                                                  uint32_t v4, uint32_t v5, uint32_t v6, uint32_t v7, long long voff,
                                                  unsigned hitmask, unsigned lane)
{
    const uint32_t vw[8] = {v0, v1, v2, v3, v4, v5, v6, v7};
    const uint2 *table = reinterpret_cast<const uint2 *>(a.aux + kBloomWords);
    unsigned nver = 0;
    unsigned pending = __ballot_sync(kFull, hitmask != 0);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        uint32_t s[8];
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = __shfl_sync(kFull, vw[i], src);
        const long long soff = __shfl_sync(kFull, voff, src);
        unsigned blocks = __shfl_sync(kFull, hitmask, src);  // bit k: block k of the vector hit the bitmap
        while (blocks) {
            const int blk = __ffs(blocks) - 1;
            blocks &= blocks - 1;
            uint32_t h;
            if (B == 4) {
                h = s[0];
#pragma unroll
                for (int i = 1; i < 8; i++) h = (blk == i) ? s[i] : h;
            } else if (B == 8) {
                h = fp8(s[0], s[1]);
#pragma unroll
                for (int i = 1; i < 4; i++) h = (blk == i) ? fp8(s[2 * i], s[2 * i + 1]) : h;
            } else {
                h = blk ? fp16(s[4], s[5], s[6], s[7]) : fp16(s[0], s[1], s[2], s[3]);
            }
            const long long t = soff + blk * B;
            const uint32_t base = table_slot(h);
            for (int round = 0; round < kSkipTableSize / 32; round++) {
                const uint2 e = __ldg(table + ((base + round * 32 + lane) & (kSkipTableSize - 1)));
                const unsigned empty = __ballot_sync(kFull, e.y == 0);
                unsigned eq = __ballot_sync(kFull, e.y != 0 && e.x == h);
                if (empty) eq &= (1u << (__ffs(empty) - 1)) - 1u;
                while (eq) {
                    const int l = __ffs(eq) - 1;
                    eq &= eq - 1;
                    const uint32_t code = __shfl_sync(kFull, e.y, l) - 1;
                    const int d = (int)(code & 31), pid = (int)(code >> 5);
                    const long long j = t - d - a.anchor;
                    if (j >= 0 && j < a.hi) {
                        nver += (lane == 0);
                        if (warp_verify(a, j, lane, a.pat + (size_t)pid * a.pat_stride_words) && lane == 0)
                            atomicAdd(a.multi_counts + pid, 1ull);
                    }
                }
                if (empty) break;
            }
        }
    }
    return nver;
}

template <int B, int U>
__global__ void __launch_bounds__(kThreads) scan_multi_kernel(const __grid_constant__ ScanArgs a)
{
    __shared__ uint32_t bloom[kBloomWords];
    for (int i = threadIdx.x; i < kBloomWords; i += kThreads) bloom[i] = __ldg(a.aux + i);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * kThreads) >> 5;
    unsigned nver = 0;
    constexpr int NB = 32 / B;  // blocks per 32-byte vector

    for (unsigned long long g = warp; g < a.ngroups; g += nwarps) {
        const uint8_t *p = a.text + g * (U * 1024ull) + lane * 32;
        V8 v[U];
#pragma unroll
        for (int c = 0; c < U; c++) v[c] = ldg_stream256(p + c * 1024);
#pragma unroll
        for (int c = 0; c < U; c++) {
            unsigned hitmask = 0;
#pragma unroll
            for (int k = 0; k < NB; k++) {
                uint32_t h;
                if (B == 4) h = v[c].w[k];
                else if (B == 8) h = fp8(v[c].w[2 * k], v[c].w[2 * k + 1]);
                else h = fp16(v[c].w[4 * k], v[c].w[4 * k + 1], v[c].w[4 * k + 2], v[c].w[4 * k + 3]);
                const uint32_t s1 = bloom_slot(h), s2 = bloom_slot2(h), s3 = bloom_slot3(h);
                const uint32_t bit = (bloom[s1 >> 5] >> (s1 & 31)) & (bloom[s2 >> 5] >> (s2 & 31)) &
                                     (bloom[s3 >> 5] >> (s3 & 31)) & 1u;
                hitmask |= bit << k;
            }
            if (__any_sync(kFull, hitmask != 0)) {
                const long long voff = (long long)(g * (U * 1024ull)) + c * 1024 + lane * 32;
                nver += multi_candidates<B>(a, v[c].w[0], v[c].w[1], v[c].w[2], v[c].w[3], v[c].w[4], v[c].w[5],
                                            v[c].w[6], v[c].w[7], voff, hitmask, lane);
            }
        }
    }
    block_reduce_count(a, 0u, nver);
}

/* ===================================================================== scan_swar<M, DENSE> */

/* Patterns of M <= 4 bytes, four start positions per instruction (SIMD within a register).
 * For text word w at byte offset o and the words w_i = text[o+i .. o+i+4) (funnel shifts of two loaded words),
 *     t = (w_0 ^ splat(x[0])) | (w_1 ^ splat(x[1])) | ... | (w_{M-1} ^ splat(x[M-1]))
 * has byte k equal to zero exactly when x occurs at o + k.  ((t & 0x7f7f7f7f) + 0x7f7f7f7f) | t has bit 7 of
 * byte k set exactly when byte k of t is non-zero (no carries between bytes), so its complement & 0x80808080
 * marks the occurrences: exact, no verification.  Cost per 4 positions: (M-1) SHF + M LOP3 + LOP3 + ADD + LOP3, against
 * one SHF + (LOP3) + ISETP per POSITION for a word compare at every byte offset.
 * DENSE: occurrences are counted with POPC in registers (no position output).
 * !DENSE: the nz words of a 32-byte vector are AND-reduced and only a vector that holds an occurrence takes the
 *         per-position path (count + optional position output).
 * 256-bit loads, U = 4 in flight, the halo word of the next lane by shuffle (M > 1). */
__device__ __forceinline__ uint32_t movemask4(uint32_t z)  // bits 7,15,23,31 -> bits 0..3
{
    return ((z >> 7) * 0x10204080u) >> 28;
}

__device__ __noinline__ unsigned swar_candidates(const ScanArgs &a, unsigned mm, long long voff)
{
    unsigned cnt = 0;
    while (mm) {
        const int p = __ffs(mm) - 1;
        mm &= mm - 1;
        const long long j = voff + p;
        if (j >= a.hi) break;
        cnt++;
        emit_match(a, j);
    }
    return cnt;
}

template <int M, bool DENSE>
__global__ void __launch_bounds__(kThreads) scan_swar_kernel(const __grid_constant__ ScanArgs a)
{
    constexpr int U = 4;
    constexpr uint32_t K7 = 0x7f7f7f7fu, K8 = 0x80808080u;
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * kThreads) >> 5;
    uint32_t C[4];
#pragma unroll
    for (int i = 0; i < 4; i++) C[i] = ((a.X0 >> (8 * i)) & 0xffu) * 0x01010101u;
    const uint32_t one = (uint32_t)a.period;  // set to 1 by launch_scan
    unsigned cnt = 0;

    for (unsigned long long g = warp; g < a.ngroups; g += nwarps) {
        const uint8_t *p = a.text + g * (U * 1024ull) + lane * 32;
        V8 v[U];
#pragma unroll
        for (int c = 0; c < U; c++) v[c] = ldg_stream256(p + c * 1024);
        uint32_t gn = 0;
        if (M > 1 && lane == 0) gn = __ldg(reinterpret_cast<const uint32_t *>(a.text + (g + 1) * (U * 1024ull)));
#pragma unroll
        for (int c = 0; c < U; c++) {
            uint32_t W[9];
#pragma unroll
            for (int i = 0; i < 8; i++) W[i] = v[c].w[i];
            W[8] = 0u;
            if (M > 1) {
                const uint32_t nxt = (c + 1 < U) ? v[(c + 1 < U) ? c + 1 : c].w[0] : gn;
                W[8] = __shfl_sync(kFull, lane == 0 ? nxt : v[c].w[0], (lane + 1) & 31);
            }
            const long long voff = (long long)(g * (U * 1024ull)) + c * 1024 + lane * 32;
            if (M == 4 && !DENSE) {
                // rare 4-byte pattern: one SHF + one ISETP.EQ.OR per position is cheaper than the byte-parallel
                // test (7 against 9 ALU operations per word); the SIMD test below runs only on a hit
                bool hit = false;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    hit |= (W[k] == a.X0);
#pragma unroll
                    for (int i = 1; i < 4; i++) hit |= (__funnelshift_r(W[k], W[k + 1], 8 * i) == a.X0);
                }
                if (!hit) continue;
            }
            uint32_t z[8];  // bit 7 of byte b of z[k] set <=> occurrence at voff + 4k + b
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t t = W[k] ^ C[0];
#pragma unroll
                for (int i = 1; i < M; i++) t |= __funnelshift_r(W[k], W[k + 1], 8 * i) ^ C[i];
                // the add as IMAD (FMA pipe) - the ALU pipe is the one that saturates; `one` comes from the
                // parameter block so that ptxas cannot fold the multiply back into an ALU add
                const uint32_t s7 = (t & K7) * one + K7;
                z[k] = ~(s7 | t) & K8;
            }
            if (DENSE && voff + 32 <= a.hi) {
#pragma unroll
                for (int k = 0; k < 8; k++) cnt += __popc(z[k]);
            } else if (voff < a.hi) {
                uint32_t any = z[0];
#pragma unroll
                for (int k = 1; k < 8; k++) any |= z[k];
                if (any != 0u) {
                    unsigned mm = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) mm |= movemask4(z[k]) << (4 * k);
                    cnt += swar_candidates(a, mm, voff);
                }
            }
        }
    }
    block_reduce_count(a, cnt, 0u);
}

/* ===================================================================== position output */

/* Occurrence positions leave the scan kernels as a BITMAP over the scanned start positions (emit_match: one
 * atomicOr per occurrence, no cursor, no ordering problem).  Three small kernels turn it into the sorted list:
 * per-CTA popcounts, an exclusive scan of those (one CTA; it also advances the running output offset, so windows
 * and shards are chained on the device without a host round trip), and an ordered write that stops at `cap`.
 * This is SURVEY section 7.6's "per-tile counts -> exclusive scan -> ordered write" with the tile's content kept as
 * bits: one scan of the text, no sort, no second scan when the caller's buffer is short. */
constexpr int kCompactWords = 4096;    // bitmap words per CTA = 131072 start positions
constexpr int kCompactThreads = 256;   // 16 consecutive words per thread

__global__ void __launch_bounds__(kCompactThreads) bitmap_count_kernel(const uint32_t *bits, unsigned long long nwords,
                                                                        unsigned *block_sums)
{
    __shared__ unsigned s_sum[kCompactThreads / 32];
    const unsigned long long w0 = (unsigned long long)blockIdx.x * kCompactWords + threadIdx.x * 16ull;
    unsigned c = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (w0 + 4 * q < nwords) {  // (nwords is a multiple of 4: the bitmap is allocated in whole CTAs)
            const uint4 v = *reinterpret_cast<const uint4 *>(bits + w0 + 4 * q);
            c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int w = 0; w < kCompactThreads / 32; w++) t += s_sum[w];
        block_sums[blockIdx.x] = t;
    }
}

/* one CTA of 1024 threads: block_off[b] = *running + sum of block_sums[0..b); then *running += total */
__global__ void __launch_bounds__(1024) bitmap_scan_kernel(const unsigned *block_sums, unsigned nblocks,
                                                             unsigned long long *block_off, unsigned long long *running)
{
    __shared__ unsigned long long s_part[1024];
    const unsigned per = (nblocks + 1023) / 1024;
    const unsigned lo = threadIdx.x * per, hi = min(nblocks, lo + per);
    unsigned long long sum = 0;
    for (unsigned b = lo; b < hi; b++) sum += block_sums[b];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive scan of the per-thread sums
        const unsigned long long v = threadIdx.x >= (unsigned)o ? s_part[threadIdx.x - o] : 0ull;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    const unsigned long long base = *running;
    unsigned long long off = base + s_part[threadIdx.x] - sum;
    for (unsigned b = lo; b < hi; b++) {
        block_off[b] = off;
        off += block_sums[b];
    }
    __syncthreads();  // every thread has read *running
    if (threadIdx.x == 1023) *running = base + s_part[1023];
}

/* ordered write: bit i of the bitmap -> out[block_off + rank] = first + i, for output slots below cap */
__global__ void __launch_bounds__(kCompactThreads) bitmap_write_kernel(const uint32_t *bits, unsigned long long nwords,
                                                                        const unsigned long long *block_off,
                                                                        unsigned long long first, unsigned long long *out,
                                                                        unsigned long long cap)
{
    __shared__ unsigned s_warp[kCompactThreads / 32];
    unsigned long long off = block_off[blockIdx.x];
    if (off >= cap) return;  // (uniform) everything this CTA holds lies beyond the caller's buffer
    const unsigned long long w0 = (unsigned long long)blockIdx.x * kCompactWords + threadIdx.x * 16ull;
    uint32_t w[16];
    unsigned c = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (w0 + 4 * q < nwords) v = *reinterpret_cast<const uint4 *>(bits + w0 + 4 * q);
        w[4 * q] = v.x;
        w[4 * q + 1] = v.y;
        w[4 * q + 2] = v.z;
        w[4 * q + 3] = v.w;
        c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    // exclusive prefix of c over the CTA's threads
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(kFull, incl, o);
        if (lane >= (unsigned)o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned before = 0;
    for (unsigned k = 0; k < warp; k++) before += s_warp[k];
    off += before + incl - c;
    if (!c) return;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        uint32_t v = w[k];
        while (v) {
            const int b = __ffs(v) - 1;
            v &= v - 1;
            if (off < cap) out[off] = first + (w0 + k) * 32ull + b;
            off++;
        }
    }
}

/* ===================================================================== text statistics */

/* byte histogram of up to `nsamples` 4 KiB samples spread evenly over the shard: used once per text
 * to estimate how selective a B-byte fingerprint is (collision entropy), see choose_filter() */
__global__ void __launch_bounds__(256) histogram_kernel(const uint8_t *text, unsigned long long len,
                                                        unsigned long long nsamples, unsigned long long stride,
                                                        unsigned long long *hist)
{
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (unsigned long long s = blockIdx.x; s < nsamples; s += gridDim.x) {
        const unsigned long long off = s * stride;
        for (unsigned i = threadIdx.x * 16; i < 4096; i += 256 * 16) {
            if (off + i + 16 <= len) {
                const uint4 v = *reinterpret_cast<const uint4 *>(text + off + i);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    atomicAdd(&h[w[k] & 0xff], 1u);
                    atomicAdd(&h[(w[k] >> 8) & 0xff], 1u);
                    atomicAdd(&h[(w[k] >> 16) & 0xff], 1u);
                    atomicAdd(&h[w[k] >> 24], 1u);
                }
            }
        }
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)h[threadIdx.x]);
}

/* ===================================================================== text generator */

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

/* 8 text bytes for 64-bit word index q of stream (seed, sigma): a counter-based generator, so any
 * slice of the text can be produced independently on any device or on the host (textgen.c's
 * `rand() % sigma` convention: raw byte values 0..sigma-1). */
__host__ __device__ __forceinline__ unsigned long long gen_word(unsigned long long seed, unsigned long long q,
                                                                int sigma)
{
    unsigned long long z = mix64(seed + (q + 1) * 0x9E3779B97F4A7C15ull);
    if (sigma == 256) return z;
    unsigned long long out = 0;
    if ((sigma & (sigma - 1)) == 0) {
        const unsigned long long msk = (unsigned long long)(sigma - 1) * 0x0101010101010101ull;
        return z & msk;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        unsigned long long b = (z >> (8 * k)) & 0xff;
        out |= ((b * (unsigned)sigma) >> 8) << (8 * k);
    }
    return out;
}

/* fills 16-byte vectors [0, nvec) of dst with text bytes starting at global offset `base` (multiple of 16);
 * bytes at global offset >= n_total are written as 0 */
__global__ void __launch_bounds__(256) textgen_kernel(uint4 *dst, unsigned long long nvec, unsigned long long base,
                                                      unsigned long long n_total, unsigned long long seed, int sigma)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const unsigned long long off = base + i * 16;
        unsigned long long w0 = 0, w1 = 0;
        if (off < n_total) {
            w0 = gen_word(seed, off >> 3, sigma);
            w1 = gen_word(seed, (off >> 3) + 1, sigma);
            if (off + 16 > n_total) {  // zero the bytes past the end of the text
                const unsigned long long keep = n_total - off;  // 1..15
                if (keep <= 8) {
                    w1 = 0;
                    if (keep < 8) w0 &= (1ull << (8 * keep)) - 1;
                } else {
                    w1 &= (1ull << (8 * (keep - 8))) - 1;
                }
            }
        }
        uint4 o;
        o.x = (uint32_t)w0;
        o.y = (uint32_t)(w0 >> 32);
        o.z = (uint32_t)w1;
        o.w = (uint32_t)(w1 >> 32);
        dst[i] = o;
    }
}

/* ---- textgen parity mode: SMART's own random buffers, generated on the device ----
 * src/textgen.c:34-54 writes data/rand<sigma>/rand<sigma>.txt as byte = rand() % sigma from ONE glibc rand() stream
 * that is never seeded.  glibc's rand() is the additive feedback generator r[i] = r[i-3] + r[i-31] (mod 2^32),
 * output r[i] >> 1: linear, so the host jumps to the start of every block of kRandBlock outputs (x^P modulo
 * x^31 - x^28 - 1, api.cu) and each thread runs the recurrence for its block from those 31 words. */
constexpr int kRandBlock = 4096;

__global__ void __launch_bounds__(128) rand_textgen_kernel(uint8_t *dst, unsigned long long nbytes, const uint32_t *states,
                                                            unsigned nblocks, int sigma)
{
    const unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    uint32_t a[31];  // a[k] = the value 31-k steps back ... kept as a ring: a[(i + k) % 31]
#pragma unroll
    for (int k = 0; k < 31; k++) a[k] = states[31ull * b + k];
    const unsigned long long lo = (unsigned long long)b * kRandBlock;
    const unsigned long long hi = min(nbytes, lo + kRandBlock);
    int head = 0;  // a[head] is the oldest value = the next output
    for (unsigned long long i = lo; i < hi; i++) {
        const uint32_t out = a[head];
        dst[i] = (uint8_t)((out >> 1) % (uint32_t)sigma);
        // the value 31 steps later: a_{n+31} = a_{n+28} + a_n
        const int k28 = head + 28 >= 31 ? head + 28 - 31 : head + 28;
        a[head] = a[k28] + out;
        head = head + 1 == 31 ? 0 : head + 1;
    }
}

}  // namespace sb
