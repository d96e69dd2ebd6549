/*
 * Table-driven scans (SB_FLAG_TABLE_SCAN): the reference algorithm's OWN preprocessing table, built bit-identically
 * by prep.c, is uploaded once per search, staged in shared memory by every CTA, and drives the scan - the literal
 * reading of "each algorithm's tables are built once and placed in constant or shared memory".
 *
 *   kmp / mp     kmpNext[0..m] / mpNext[0..m] (kmp.c:27-41, mp.c:26-35) + the pattern bytes: the automaton loop of
 *                kmp.c:56-65, one text byte per step, failure links followed in shared memory
 *   hor          hbc[256] (hor.c:26-30): window compare left to right, then s += hbc[y[s+m-1]] (hor.c:43-48)
 *   so / sa      S[256] and lim / F over the first min(m,32) characters (so.c:27-38, sa.c:27-34): D = (D<<1)|S[y[j]],
 *                occurrence iff D < lim (so.c:54-58); m > 32: the 32-character prefix automaton + verification of the
 *                remaining bytes, as search_large does (so.c:85-93) - without reading past the text
 *   hash3/5/8    shift[256] and sh1 (hash3.c:37-55): skip while shift[h(q-gram)] != 0, verify at 0, then += sh1
 *                (hash3.c:66-81) - without the pattern sentinel the reference writes behind the text (hash3.c:61)
 *   bm           bmBc[256] + bmGs[0..m) (bm.c:27-66): right-to-left compare, j += MAX(bmGs[i], bmBc[y[i+j]]-m+1+i)
 *   qs           qsBc[256] (qs.c:26-30): compare, then s += qsBc[y[s+m]] (qs.c:46-50)
 *   raita        bmBc[256] (raita.c:26-32): last, middle, first character, then the rest (raita.c:54-60)
 *   tunedbm      bmBc[256] with bmBc[x[m-1]] = 0 and `shift` (tunedbm.c:38-42): the unrolled skip loop (:48-53)
 *                bounded by the segment instead of the sentinel the reference writes behind the text (:42)
 *   kr           d = 2^(m-1) and hx (kr.c:34-40): rolling hash REHASH (kr.c:26), memcmp on a hash hit (:48)
 *   qf<Q><S>     B[1 << (Q*S)] (qf23.c:50-56): the q-gram at every (m-Q+1)-th position and, while bits survive,
 *                every Q-th one before it; surviving alignments are compared (qf23.c:61-82).  B sits in shared
 *                memory up to 8192 entries, else in global memory (L2: 2^16 or 2^18 entries)
 *   aut          delta[(m+1)*256] (aut.c:26-45, preSMA): state = delta[state][y[j]], occurrence iff state == m; 16-bit
 *                entries, in shared memory up to m = 94 (47.5 KiB), else read through L1/L2
 *   bom          the factor oracle of the reversed pattern as a dense table trans[(m+1)*256] + terminal flags
 *                (bom.c:49-70): windows are read right to left through the oracle, the shift is the position of the
 *                last terminal state seen (bom.c:90-106); shared memory up to m = 93, else L1/L2
 *   ebom         the same oracle + the first-transition table FT[256][256] = state after the window's last two
 *                characters (ebom.c:54-59): windows without a defined FT entry are skipped m-1 at a time (ebom.c:68);
 *                the oracle rows through L1/L2, FT (128 KiB) likewise
 *   sbndm2       (m <= 32; the reference's default selection, algorithms.h:133) B[256] = the pattern's positions per
 *                character, read right to left; D = (B[y[j]] << 1) & B[y[j-1]] on the window's last two characters,
 *                then one more character per step while D != 0 (sbndm2.c:52-66)
 *   bndm         B[256] (bndm.c:33-38; the 32-character prefix for m > 32): the window is read right to left through
 *                the bit-parallel suffix automaton, `last` = the longest window suffix that is a pattern prefix
 *                (bndm.c:44-58); m > 32: search_large's prefix automaton + comparison of the rest (bndm.c:62-110)
 *   br / tvsbs   brBc[256][256] (br.c:26-37): compare (tvsbs: last and first character first, tvsbs.c:50-56), then
 *                j += brBc[y[j+m]][y[j+m+1]] (br.c:58); the 256 KiB table is read through L1/L2
 *   smith        bmBc[256] + qsBc[256] (smith.c:41-63): j += MAX(bmBc[y[j+m-1]], qsBc[y[j+m]])
 *   skip         z[256] + list[m] (skip.c:31-39): every m-th text character selects the alignments to compare
 *   fjs          qsBc[256] + kmpNext[0..m] (fjs.c:47-73): Sunday shifts until the last character matches, then a
 *                left-to-right compare and a KMP shift
 *   ssabs        qsBc[256] (ssabs.c:35-58): last and first character, the inside backwards, then j += qsBc[y[j+m]]
 *   fs           bc[256] over all m characters + Pre_GS (fs.c:26-51): skip until bc == 0, compare backwards, shift gs
 *   graspm       bc[256] + candidate lists keyed by the character before (graspm.c:37-70)
 *   tbm          bmBc[256] + bmGs[0..m) (tbm.c:66-111): Turbo-BM, the factor matched by the previous attempt is
 *                remembered (u) and jumped over; a segment starts with u = 0 like the text does
 *   ac           kmpNext[0..m] + ell (ac.c:44-86): Apostolico-Crochemore
 *   bndmq2/4/6   B[256] (bndm.c:33-38): the window's last q characters at once, then right to left (bndmq2.c:59-73)
 *   fndm         the Shift-Or masks driving a forward factor scan from the window's last character (fndm.c:44-57)
 *   bxs, bxs2..8 B[256] (bndm.c:33-38): BNDM over every (m-q+1)-th q-gram with the sample grid re-anchored at prefixes
 *   lbndm        B[256] with ceil(m/32) pattern characters per bit (lbndm.c:53-61)
 *   sbndm-bmh    sbndm + hbc[256] = 2m-1-i (sbndm-bmh.c:41-42): a Horspool-like skip in front of every window
 *   sbndm        B[256] aligned to the top of the word + the pattern's period as the shift (sbndm.c:44-83)
 *   iom          ibc1[256], ibc2[256] and the two look-ahead distances (iom.c:28-41)
 *   qlqs         qsBc of the pattern and of its reverse: the Quick-Search shift or a leap of m + m/2 (qlqs.c:52-57)
 *   tsw, tvsbs-w2  brBc of the pattern and of its reverse: one window moves up, one moves down (tsw.c:58-68)
 *   fbom, sfbom  the oracle of bom + a look-ahead FT[256][256] over the character behind the window (fbom.c:53-64)
 *   wom          wo[256] for the worst occurrence position under the character frequencies of the text's first 100
 *                characters (wom.c:28-56); the scan loop is wc's (wom.c:70-76)
 *   pbmh         hbc[256] + the comparison order by those frequencies (pbmh.c:36-47)
 *   svm0         cv[256] (svm0.c:44-52): the shift vector of the next 32 window positions
 *   wc           bc[256] for the worst character position (wc.c:28-70)
 *   fs-w1        fs's tables (fs-w1.c:72-80)
 *   aoso2/4/6    B[256] of q interleaved Shift-Or automata and their top-bit mask mm (aoso2.c:50-62)
 *   bsdm, bsdm4/5/7/8  pos[65536] over the pattern's longest run of distinct q-grams (bsdm.c:33-55)
 *   ts           no table: Tailed-Substring's two probing phases over the pattern in shared memory (ts.c:25-59)
 *   kmpskip      mpNext, kmpNext and the character buckets z[256] / list[m] (kmpskip.c:59-146), 16-bit entries
 *   zt           ztBc[256][256] (zt.c:64-74, 256 KiB through L1/L2) + bmGs[0..m): BM with a two-character shift
 *   skip3/5..8   q-gram buckets head[65536] + next[m] (skip5.c:36-45): every (m-Q+1)-th text q-gram selects the
 *                alignments to compare (skip5.c:50-57)
 *   wfr*, wfrq*, twfr*, twfrq*  the weak-factor bit table F (wfr.c:41-58, wfrq4.c:50-65; 65536 bits = 8 KiB of shared memory):
 *                the window is hashed right to left while the hash is some pattern factor's (wfr.c:73-92)
 *
 * Parallelisation: the scan loops are inherently sequential, so the text is cut into segments of `seg` start
 * positions, one THREAD per segment, each reading m-1 bytes past its segment (the same m-1 halo rule as the shards).
 * A thread that starts its automaton at the segment's first byte finds exactly the occurrences that start inside
 * the segment.  Lanes of a warp therefore read addresses `seg` bytes apart: every load instruction touches 32
 * sectors, and the kernels that read single bytes are bound by the L1/LSU wavefront rate, not by HBM (the forward
 * automata stream 256-bit vectors instead, see stream_segment) - which is why these are a selectable mode and not
 * the default.
 * What they return is bit-identical to everything else: the count of bf.c:33-36.
 *
 * Reads behind a window.  Some shift rules look at text behind the window they have just compared (qs, smith, ssabs,
 * wc, wom, fbom: one character; br, tvsbs, tsw, tvsbs-w2: two; qlqs: m/2; iom: up to m).  A shard keeps
 * SB_XSIZE-1 bytes of halo behind the start positions it owns, so such a read can land in the zero padding (last shard:
 * behind the text; other shards, m near the maximum: behind the halo).  That is harmless because every one of these
 * rules uses the character at a position p only to rule out windows that CONTAIN p: with p >= own + SB_XSIZE-1 those
 * windows start at p-m+1 >= own or later - not this shard's - and behind the text they are no windows at all.  The
 * same holds on the left for the two-ended scans (a window containing a position before local byte 0 is not owned).
 */
#pragma once
#include "scan_kernels.cuh"

namespace sb {

struct TableArgs {
    const uint8_t *text;          // shard buffer (same padding rules as ScanArgs::text)
    const uint8_t *pat;           // pattern bytes (device), m of them
    const int32_t *table;         // device copy of the algorithm's table (see each kernel)
    unsigned long long *counters; // [0] match count
    long long hi;                 // owned start positions [0, hi)
    int m;
    int seg;                      // start positions per thread (multiple of 64: the streaming kernels load aligned vectors)
    int q, sh1;                   // HASHq: q-gram length, the shift after a verification (hash3.c:53-55)
    unsigned lim;                 // Shift-Or: lim (so.c:37); Shift-And: F (sa.c:45)
};

constexpr int kTableThreads = 128;

__device__ __forceinline__ void table_reduce(const TableArgs &a, unsigned cnt)
{
    __shared__ unsigned s_cnt[kTableThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long c = 0;
#pragma unroll
        for (int w = 0; w < kTableThreads / 32; w++) c += s_cnt[w];
        if (c) atomicAdd(a.counters, c);
    }
}

/*
 * Read-only view of a shared-memory table.  nvcc for sm_100a re-derives the shared window base (S2R SR_CgaCtaId +
 * LEA) in front of every indexed LDS inside these branchy loops - a 20+ cycle special-register read on the
 * dependency chain of every table lookup.  The view converts the base once, launders it through a volatile move
 * so it cannot be rematerialised, and reads with ld.shared at base + index * sizeof(T).  Constructed after the
 * __syncthreads() that ends the staging; the tables are never written again.
 */
template <class T>
struct Smem {
    uint32_t base;
    __device__ __forceinline__ explicit Smem(const T *p)
    {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
        asm volatile("mov.u32 %0, %1;" : "=r"(base) : "r"(a));
    }
    __device__ __forceinline__ T at_byte(uint32_t byte_off) const  // byte_off a multiple of sizeof(T)
    {
        uint32_t v;
        if (sizeof(T) == 1)
            asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + byte_off));
        else if (sizeof(T) == 2)
            asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(base + byte_off));
        else
            asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + byte_off));
        return (T)v;
    }
    __device__ __forceinline__ T operator[](int i) const { return at_byte((uint32_t)i * (uint32_t)sizeof(T)); }
};

/* pattern bytes into shared memory (padded to a word) */
__device__ __forceinline__ void stage_pattern_smem(uint8_t *sx, const TableArgs &a)
{
    for (int i = threadIdx.x; i < a.m; i += kTableThreads) sx[i] = a.pat[i];
}

/*
 * The forward automata (kmp, mp, so, sa, aut) consume the text strictly one byte after the other, so a thread streams
 * its segment with 256-bit loads - 64 B (two sectors of one line, a full DRAM burst) per step, the next 64 B
 * requested before the current ones are consumed - and takes the bytes out of registers at compile-time offsets.
 * A warp-wide byte or word load over 32 segments costs 32 L1 wavefronts per 32 / 128 bytes; the vector load pays
 * the same 32 wavefronts per 1024 bytes, which moves these kernels from the LSU bound to their shared-memory
 * table lookups.  `lo` is a multiple of 64 (segments are multiples of 64 start positions); the last step reads up to 127 bytes
 * past `end`, inside the shard buffer's zeroed back pad (kPadBytes).  f(c, j): byte c at text index j.
 */
struct StreamNoFlush {
    __device__ __forceinline__ void operator()(long long) const {}
};

template <class F, class G>
__device__ __forceinline__ void stream_words(const V8 &v, long long j0, F &&f, G &&flush)
{
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t w = v.w[k];
        f(w & 0xffu, j0 + 4 * k, 4 * k);
        f((w >> 8) & 0xffu, j0 + 4 * k + 1, 4 * k + 1);
        f((w >> 16) & 0xffu, j0 + 4 * k + 2, 4 * k + 2);
        f(w >> 24, j0 + 4 * k + 3, 4 * k + 3);
    }
    flush(j0);
}

/* f(c, j, k): byte c at text index j, k = j - j0 in 0..31 a compile-time constant; flush(j0) after each 32 bytes */
template <class F, class G = StreamNoFlush>
__device__ __forceinline__ void stream_segment(const uint8_t *text, long long lo, long long end, F &&f, G &&flush = G())
{
    V8 a = ldg_stream256(text + lo), b = ldg_stream256(text + lo + 32);
    long long j = lo;
    for (; j + 64 <= end; j += 64) {
        const V8 na = ldg_stream256(text + j + 64), nb = ldg_stream256(text + j + 96);
        stream_words(a, j, f, flush);
        stream_words(b, j + 32, f, flush);
        a = na;
        b = nb;
    }
    if (j < end) {  // the last, partial step
        auto g = [&](unsigned c, long long jj, int k) {
            if (jj < end) f(c, jj, k);
        };
        stream_words(a, j, g, flush);
        stream_words(b, j + 32, g, flush);
    }
}

/* ---- Knuth-Morris-Pratt / Morris-Pratt: table = kmpNext[0..m] or mpNext[0..m] ---- */
__global__ void __launch_bounds__(kTableThreads) table_kmp_kernel(const __grid_constant__ TableArgs a)
{
    // one shared-memory word per pattern position: x[i] in the high half, kmpNext[i] (>= -1) in the low half, so a
    // step of the failure loop is ONE lookup on the dependency chain instead of two
    __shared__ unsigned raw_xn[SB_XSIZE + 1];
    for (int i = threadIdx.x; i < a.m; i += kTableThreads)
        raw_xn[i] = ((unsigned)a.pat[i] << 16) | ((unsigned)a.table[i] & 0xffffu);
    __syncthreads();
    const Smem<unsigned> s_xn(raw_xn);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);   // starts [lo, own_end)
        const long long end = own_end + a.m - 1;           // text bytes [lo, end)
        int i = 0;
        const int m = a.m, next_m = __ldg(a.table + a.m);
        stream_segment(a.text, lo, end, [&](unsigned c, long long, int) {
            while (i > -1) {                                // kmp.c:57-58: while (i > -1 && x[i] != y[j]) i = kmpNext[i]
                const unsigned v = s_xn[i];
                if ((v >> 16) == c) break;
                i = (int)(short)(v & 0xffffu);
            }
            i++;
            if (i >= m) {                                   // kmp.c:61-64: OUTPUT(j - i) with j already advanced
                cnt++;                                      // the start j+1-m is >= lo (i began at 0) and < own_end
                i = next_m;
            }
        });
    }
    table_reduce(a, cnt);
}

/* ---- Horspool: table = hbc[256] ---- */
__global__ void __launch_bounds__(kTableThreads) table_hor_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_hbc[256];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_hbc[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_hbc(raw_hbc);
    const Smem<uint8_t> s_x(raw_x);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long s = lo;
        while (s < own_end) {
            int i = 0;
            while (i < a.m && s_x[i] == __ldg(a.text + s + i)) i++;  // hor.c:45-46
            if (i == a.m) cnt++;
            s += s_hbc[__ldg(a.text + s + a.m - 1)];                 // hor.c:48
        }
    }
    table_reduce(a, cnt);
}

/* ---- Shift-Or (AND = false) / Shift-And (AND = true): table = S[256] over the first min(m,32) characters ---- */
/* so.c:88-91 / sa.c:86-89: the bytes after the 32-character prefix (m > 32 only; out of line, it is rare) */
__device__ __noinline__ unsigned so_verify_rest(const uint8_t *text, const uint8_t *pat, long long st, int m)
{
    int i = SB_WORD;
    while (i < m && __ldg(pat + i) == __ldg(text + st + i)) i++;
    return i == m;
}

template <bool AND, bool LARGE>
__global__ void __launch_bounds__(kTableThreads) table_so_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_S[256 * 32];  // one copy of S per lane, entry c of lane l in bank l: lookups never conflict
    for (int i = threadIdx.x; i < 256 * 32; i += kTableThreads) raw_S[i] = (unsigned)a.table[i >> 5];
    __syncthreads();
    const Smem<unsigned> S(raw_S + (threadIdx.x & 31));
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    const int k = LARGE ? SB_WORD : a.m;  // characters in the automaton (search_large: the 32-character prefix)
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        const long long end = own_end + k - 1;  // the prefix automaton reads [lo, end)
        unsigned D = AND ? 0u : ~0u;
        const unsigned lim = a.lim;
        unsigned hits = 0;  // LARGE: prefix hits among the 32 bytes in flight, verified after them
        stream_segment(
            a.text, lo, end,
            [&](unsigned c, long long, int b) {
                bool hit;
                if (AND) {
                    D = ((D << 1) | 1u) & S.at_byte(c * 128);  // sa.c:50
                    hit = (D & lim) != 0;                      // sa.c:51 (F = 1 << (k-1))
                } else {
                    D = (D << 1) | S.at_byte(c * 128);         // so.c:55
                    hit = D < lim;                             // so.c:56
                }
                if (LARGE) {
                    if (hit) hits |= 1u << b;
                } else {
                    cnt += hit;
                }
            },
            [&](long long j0) {
                if (!LARGE) return;
                while (hits) {  // the start j - 31 is >= lo: D began empty at lo
                    const int b = __ffs(hits) - 1;
                    hits &= hits - 1;
                    cnt += so_verify_rest(a.text, a.pat, j0 + b - (SB_WORD - 1), a.m);
                }
            });
    }
    table_reduce(a, cnt);
}

/* ---- HASHq: table = shift[256]; a.q = 3, 5 or 8; a.sh1 ---- */
__global__ void __launch_bounds__(kTableThreads) table_hashq_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_shift[256];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_shift[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_shift(raw_shift);
    const Smem<uint8_t> s_x(raw_x);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long i = lo + a.m - 1;  // index of the window's last byte (hash3.c:63)
        while (i - (a.m - 1) < own_end) {
            unsigned h = 0;          // hash3.c:66-68, hash5/hash8: ((..(y[i-q+1] << 1) + ..) << 1) + y[i], 8 bits
            for (int t = a.q - 1; t >= 0; t--) h = (h << 1) + __ldg(a.text + i - t);
            const int sh = s_shift[h & 0xffu];
            if (sh != 0) {
                i += sh;
                continue;
            }
            const long long st = i - (a.m - 1);  // hash3.c:73-78
            int j = 0;
            while (j < a.m && s_x[j] == __ldg(a.text + st + j)) j++;
            if (j == a.m) cnt++;
            i += a.sh1;
        }
    }
    table_reduce(a, cnt);
}

/* ---- Boyer-Moore: table = bmBc[256] followed by bmGs[0..m) ---- */
__global__ void __launch_bounds__(kTableThreads) table_bm_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_bc[256];
    __shared__ int raw_gs[SB_XSIZE];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_bc[i] = a.table[i];
    for (int i = threadIdx.x; i < a.m; i += kTableThreads) raw_gs[i] = a.table[256 + i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_bc(raw_bc);
    const Smem<int> s_gs(raw_gs);
    const Smem<uint8_t> s_x(raw_x);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long j = lo;
        while (j < own_end) {
            int i = a.m - 1;
            unsigned c = 0;
            while (i >= 0 && s_x[i] == (c = __ldg(a.text + i + j))) --i;  // bm.c:83
            if (i < 0) {
                cnt++;
                j += s_gs[0];                                              // bm.c:86
            } else {
                const int bc = s_bc[c] - a.m + 1 + i;
                j += s_gs[i] > bc ? s_gs[i] : bc;                          // bm.c:89
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- bad-character scans: table = qsBc[256] (QS) or bmBc[256] (RAITA, TUNED) ---- */
enum { kBcQs = 0, kBcRaita = 1, kBcTuned = 2 };
template <int KIND>
__global__ void __launch_bounds__(kTableThreads) table_bc_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_bc[256];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_bc[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const int m = a.m;
    const unsigned last_ch = raw_x[m - 1];
    const int shift = raw_bc[last_ch];                  // tunedbm.c:40 (read before the entry is zeroed)
    __syncthreads();
    if (KIND == kBcTuned && threadIdx.x == 0) raw_bc[last_ch] = 0;  // tunedbm.c:41
    __syncthreads();
    const Smem<int> s_bc(raw_bc);
    const Smem<uint8_t> s_x(raw_x);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long s = lo;
        while (s < own_end) {
            if (KIND == kBcQs) {
                int i = 0;
                while (i < m && s_x[i] == __ldg(a.text + s + i)) i++;  // qs.c:47-48
                if (i == m) cnt++;
                s += s_bc[__ldg(a.text + s + m)];                      // qs.c:50 (the byte after the window)
            } else if (KIND == kBcRaita) {
                const unsigned c = __ldg(a.text + s + m - 1);
                if (c == last_ch && s_x[m / 2] == __ldg(a.text + s + m / 2) && s_x[0] == __ldg(a.text + s)) {  // raita.c:55-57
                    int i = 1;
                    while (i < m - 1 && s_x[i] == __ldg(a.text + s + i)) i++;                                   // raita.c:58
                    if (i >= m - 1) cnt++;
                }
                s += s_bc[c];                                                                                  // raita.c:60
            } else {
                int k = s_bc[__ldg(a.text + s + m - 1)];  // tunedbm.c:48-53, bounded by the segment, not a sentinel
                while (k != 0 && s < own_end) {
                    s += k;
                    k = s_bc[__ldg(a.text + s + m - 1)];
                }
                if (s < own_end && k == 0) {
                    int i = 0;
                    while (i < m - 1 && s_x[i] == __ldg(a.text + s + i)) i++;  // tunedbm.c:54: the first m-1 bytes
                    if (i >= m - 1) cnt++;
                }
                s += shift;                                                     // tunedbm.c:56
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Karp-Rabin: a.lim = d = 2^(m-1) (wrapping), a.sh1 = hx ---- */
/*
 * kr.c:44-52 is a forward stream with two read heads m bytes apart: hy = REHASH(y[j], y[j+m], hy) =
 * ((hy - y[j]*d) << 1) + y[j+m] (kr.c:26).  The thread streams the LEADING head with 256-bit loads like the other
 * forward automata; the trailing byte y[e-m] of step e comes out of the previous and the current vector, shifted by
 * 32-m bytes once per 32 bytes (m <= 32).  For m >= 33 the reference's int d = 2^(m-1) has wrapped to 0 (kr.c:38)
 * and the trailing head drops out of the arithmetic altogether (LAG = false).  The first window of a segment is the
 * same recurrence with nothing to take out (kr.c:39-41), so the trailing bytes before the segment read as 0.
 */
__device__ __noinline__ unsigned kr_verify(const uint8_t *text, const uint8_t *pat, long long st, int m)
{
    int i = 0;
    while (i < m && __ldg(pat + i) == __ldg(text + st + i)) i++;  // memcmp, kr.c:48
    return i == m;
}

template <bool LAG>
__global__ void __launch_bounds__(kTableThreads) table_kr_kernel(const __grid_constant__ TableArgs a)
{
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const int m = a.m;
        const long long own_end = min(lo + a.seg, a.hi);
        const long long end = own_end + m - 1;  // the leading head reads [lo, end)
        const unsigned negd = 0u - a.lim, hx = (unsigned)a.sh1;
        const int q = LAG ? (32 - m) >> 2 : 0;
        const unsigned r8 = LAG ? 8u * ((32 - m) & 3) : 0u;
        unsigned hy = 0;
        V8 P = {{0, 0, 0, 0, 0, 0, 0, 0}};
        V8 C = ldg_stream256(a.text + lo);
        for (long long j0 = lo; j0 < end; j0 += 32) {
            const V8 N = ldg_stream256(a.text + j0 + 32);  // (at most 63 bytes past `end`: the back pad)
            V8 L = {{0, 0, 0, 0, 0, 0, 0, 0}};
            if (LAG) {  // L = the 32 bytes that start m before j0
                uint32_t X[16];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    X[i] = P.w[i];
                    X[8 + i] = C.w[i];
                }
#define SB_KR_LAG(Q)                                                                      \
    case Q:                                                                               \
        _Pragma("unroll") for (int i = 0; i < 8; i++) L.w[i] = __funnelshift_r(X[Q + i], X[Q + i + 1], r8); \
        break;
                switch (q) {
                    SB_KR_LAG(0) SB_KR_LAG(1) SB_KR_LAG(2) SB_KR_LAG(3) SB_KR_LAG(4) SB_KR_LAG(5) SB_KR_LAG(6)
                    default:
#pragma unroll
                        for (int i = 0; i < 8; i++) L.w[i] = __funnelshift_r(X[7 + i], X[8 + i], r8);
                        break;
                }
#undef SB_KR_LAG
            }
            unsigned hits = 0;
#pragma unroll
            for (int k = 0; k < 32; k++) {
                const unsigned c = (C.w[k >> 2] >> (8 * (k & 3))) & 0xffu;
                if (LAG) {
                    const unsigned l = (L.w[k >> 2] >> (8 * (k & 3))) & 0xffu;
                    hy = ((hy + l * negd) << 1) + c;  // REHASH, kr.c:26,49
                } else {
                    hy = (hy << 1) + c;               // d == 0
                }
                if (hy == hx) hits |= 1u << k;        // kr.c:48, the comparison after these 32 bytes
            }
            while (hits) {
                const int k = __ffs(hits) - 1;
                hits &= hits - 1;
                const long long st = j0 + k - m + 1;  // the window that ends at byte j0 + k
                if (st >= lo && st < own_end) cnt += kr_verify(a.text, a.pat, st, m);
            }
            P = C;
            C = N;
        }
    }
    table_reduce(a, cnt);
}

/* ---- deterministic automaton: table = delta as 16-bit entries, (m+1) rows of 256 ---- */
/* the rows sit in DYNAMIC shared memory, (m+1) * 512 bytes, when that is at most kAutSharedBytes (four CTAs per SM
 * still fit); longer patterns read delta through L1/L2 (smem == nullptr view: `rows` = 0) */
constexpr int kAutSharedBytes = 56 * 1024;
__global__ void __launch_bounds__(kTableThreads) table_aut_kernel(const __grid_constant__ TableArgs a)
{
    extern __shared__ __align__(16) unsigned short raw_delta[];
    const unsigned short *gdelta = reinterpret_cast<const unsigned short *>(a.table);
    const bool small = (a.m + 1) * 512 <= kAutSharedBytes;
    if (small) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(gdelta);
        uint32_t *dst = reinterpret_cast<uint32_t *>(raw_delta);
        for (int i = threadIdx.x; i < (a.m + 1) * 128; i += kTableThreads) dst[i] = src[i];
    }
    __syncthreads();
    const Smem<unsigned short> s_delta(raw_delta);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        const long long end = own_end + a.m - 1;
        unsigned state = 0;  // a segment starts like the text does: only occurrences that start at or after lo are seen
        const unsigned m = (unsigned)a.m;
        if (small) {
            stream_segment(a.text, lo, end, [&](unsigned c, long long, int) {
                state = s_delta.at_byte(state * 512 + c * 2);  // aut.c:39
                cnt += state == m;                             // aut.c:40
            });
        } else {
            stream_segment(a.text, lo, end, [&](unsigned c, long long, int) {
                state = __ldg(gdelta + state * 256 + c);
                cnt += state == m;
            });
        }
    }
    table_reduce(a, cnt);
}

/* ---- Backward Oracle Matching: table = trans as 16-bit entries ((m+1) rows of 256), then m+1 terminal bytes ---- */
constexpr int kBomSharedRows = 94;
template <bool BOM2>  // bom2.c:58-75: the same oracle without the terminal flags, shift = what was left unread (+1 on a match)
__global__ void __launch_bounds__(kTableThreads) table_bom_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ short raw_trans[kBomSharedRows * 256];
    __shared__ unsigned char raw_term[kBomSharedRows];
    const int m = a.m;
    const short *gtrans = reinterpret_cast<const short *>(a.table);
    const unsigned char *gterm = reinterpret_cast<const unsigned char *>(gtrans + (size_t)(m + 1) * 256);
    const bool small = m + 1 <= kBomSharedRows;
    if (small) {
        for (int i = threadIdx.x; i < (m + 1) * 256; i += kTableThreads) raw_trans[i] = gtrans[i];
        for (int i = threadIdx.x; i <= m; i += kTableThreads) raw_term[i] = gterm[i];
    }
    __syncthreads();
    const Smem<short> s_trans(raw_trans);
    const Smem<unsigned char> s_term(raw_term);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long j = lo;
        while (j < own_end) {       // bom.c:90-106
            int i = m - 1, r = m, shift = m, period = m, q;
            while (i >= 0 && (q = small ? s_trans[r * 256 + __ldg(a.text + i + j)] : __ldg(gtrans + r * 256 + __ldg(a.text + i + j))) >= 0) {
                r = q;
                if (!BOM2 && (small ? s_term[r] : __ldg(gterm + r))) {
                    period = shift;
                    shift = i;
                }
                --i;
            }
            if (BOM2) {           // bom2.c:69-73
                cnt += i < 0;
                j += i < 0 ? 1 : i + 1;
                continue;
            }
            if (i < 0) {
                cnt++;
                shift = period;
            }
            j += shift;
        }
    }
    table_reduce(a, cnt);
}

/* ---- Extended BOM: table = trans ((m+1) x 256 x 16 bit), m+1 terminal bytes (unused), padding to 4 bytes, then
 *      FT[256][256] as 16-bit entries (FT[a][b] = trans[trans[m][a]][b], ebom.c:54-59); a.lim = offset of FT in shorts ---- */
__global__ void __launch_bounds__(kTableThreads) table_ebom_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const short *trans = reinterpret_cast<const short *>(a.table);
    const short *FT = trans + a.lim;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        {   // ebom.c:64: the first window of the segment is compared directly
            int i = 0;
            while (i < m && s_x[i] == __ldg(a.text + lo + i)) i++;
            cnt += i == m;
        }
        long long j = lo + m;  // index of the window's last character (ebom.c:65)
        while (j - m + 1 < own_end) {
            int p;
            while ((p = __ldg(FT + __ldg(a.text + j) * 256 + __ldg(a.text + j - 1))) < 0) {  // ebom.c:68
                j += m - 1;
                if (j - m + 1 >= own_end) break;  // (the reference stops on the pattern copy it writes behind the text)
            }
            if (j - m + 1 >= own_end) break;
            long long i = j - 2;
            while (i >= j - m && (p = __ldg(trans + p * 256 + __ldg(a.text + i))) >= 0) i--;  // ebom.c:71 (never past the
            if (i < j - (m - 1)) {                                                                // window: state 0 has no exits)
                cnt++;   // ebom.c:72-75
                i++;
            }
            j = i + m;
        }
    }
    table_reduce(a, cnt);
}

/* ---- SBNDM2 (m <= 32): table = B[256]; a.sh1 = the shift after an occurrence ---- */
__global__ void __launch_bounds__(kTableThreads) table_sbndm2_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_B[256];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_B[i] = (unsigned)a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<unsigned> s_B(raw_B);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        {   // sbndm2.c:51: the first window is compared directly
            int i = 0;
            while (i < m && s_x[i] == __ldg(a.text + lo + i)) i++;
            cnt += i == m;
        }
        long long j = lo + m;  // index of the window's last character
        while (j - m + 1 < own_end) {
            unsigned D = (s_B[__ldg(a.text + j)] << 1) & s_B[__ldg(a.text + j - 1)];  // sbndm2.c:54
            if (D != 0) {
                const long long pos = j;
                while ((D = (D << 1) & s_B[__ldg(a.text + j - 2)]) != 0) --j;           // sbndm2.c:57-58
                j += m - 2;
                if (j == pos) {                                                         // sbndm2.c:60-63
                    cnt++;
                    j += a.sh1;
                }
            } else {
                j += m - 1;                                                             // sbndm2.c:65
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- BNDM: table = B[256] over the first min(m,32) characters, reversed ---- */
__global__ void __launch_bounds__(kTableThreads) table_bndm_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_B[256];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_B[i] = (unsigned)a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<unsigned> s_B(raw_B);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m, w = m < SB_WORD ? m : SB_WORD;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long j = lo;
        while (j < own_end) {  // bndm.c:44-58
            int i = w - 1, last = w;
            unsigned D = ~0u;
            while (i >= 0 && D != 0) {
                D &= s_B[__ldg(a.text + j + i)];
                i--;
                if (D != 0) {
                    if (i >= 0) {
                        last = i + 1;
                    } else {  // the (prefix of the) pattern matched: bndm.c:52 / search_large :97-101
                        int k = w;
                        while (k < m && s_x[k] == __ldg(a.text + j + k)) k++;
                        cnt += k == m;
                    }
                }
                D <<= 1;
            }
            j += last;
        }
    }
    table_reduce(a, cnt);
}

/* ---- Berry-Ravindran (TVSBS = false) / TVSBS (true): table = brBc[256*256] ---- */
template <bool TVSBS>
__global__ void __launch_bounds__(kTableThreads) table_br_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long j = lo;
        while (j < own_end) {
            if (!TVSBS) {  // br.c:54-56
                int i = 0;
                while (i < m && s_x[i] == __ldg(a.text + j + i)) i++;
                cnt += i >= m;
            } else if (s_x[m - 1] == __ldg(a.text + j + m - 1) && s_x[0] == __ldg(a.text + j)) {  // tvsbs.c:50-56
                int i = m - 2;
                while (i > 0 && s_x[i] == __ldg(a.text + j + i)) i--;
                cnt += i <= 0;
            }
            // the two characters after the window (behind the text they read as 0, like the driver's NUL bytes)
            j += __ldg(a.table + __ldg(a.text + j + m) * 256 + __ldg(a.text + j + m + 1));  // br.c:58
        }
    }
    table_reduce(a, cnt);
}

/* ---- Smith: table = bmBc[256] then qsBc[256] ---- */
__global__ void __launch_bounds__(kTableThreads) table_smith_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_bc[512];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 512; i += kTableThreads) raw_bc[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_bc(raw_bc);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long j = lo;
        while (j < own_end) {  // smith.c:54-60
            int i = 0;
            while (i < m && s_x[i] == __ldg(a.text + j + i)) i++;
            cnt += i == m;
            const int b1 = s_bc[__ldg(a.text + j + m - 1)], b2 = s_bc[256 + __ldg(a.text + j + m)];
            j += b1 > b2 ? b1 : b2;
        }
    }
    table_reduce(a, cnt);
}

/* ---- Skip Search: table = z[256] then list[m] ---- */
__global__ void __launch_bounds__(kTableThreads) table_skip_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_z[256];
    __shared__ int raw_list[SB_XSIZE];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_z[i] = a.table[i];
    for (int i = threadIdx.x; i < a.m; i += kTableThreads) raw_list[i] = a.table[256 + i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_z(raw_z);
    const Smem<int> s_list(raw_list);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        // skip.c:44-49: one text character out of m; an occurrence that starts in [lo, own_end) contains exactly one
        // of the sample points lo+m-1, lo+2m-1, ...
        for (long long j = lo + m - 1; j - (m - 1) < own_end; j += m)
            for (int e = s_z[__ldg(a.text + j)]; e >= 0; e = s_list[e]) {
                const long long st = j - e;
                if (st < lo || st >= own_end) continue;
                int i = 0;
                while (i < m && s_x[i] == __ldg(a.text + st + i)) i++;
                cnt += i == m;
            }
    }
    table_reduce(a, cnt);
}

/* ---- Franek-Jennings-Smyth: table = qsBc[256] then kmpNext[0..m] ---- */
__global__ void __launch_bounds__(kTableThreads) table_fjs_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_qs[256];
    __shared__ int raw_next[SB_XSIZE + 1];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_qs[i] = a.table[i];
    for (int i = threadIdx.x; i <= a.m; i += kTableThreads) raw_next[i] = a.table[256 + i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_qs(raw_qs);
    const Smem<int> s_next(raw_next);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long s = lo;
        while (s < own_end) {  // fjs.c:57-70
            while (s < own_end && s_x[m - 1] != __ldg(a.text + s + m - 1)) s += s_qs[__ldg(a.text + s + m)];
            if (s >= own_end) break;
            int i = 0;
            while (i < m && s_x[i] == __ldg(a.text + s + i)) i++;
            cnt += i >= m;
            s += i - s_next[i];
        }
    }
    table_reduce(a, cnt);
}

/* ---- SSABS: table = qsBc[256] ---- */
__global__ void __launch_bounds__(kTableThreads) table_ssabs_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_qs[256];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_qs[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_qs(raw_qs);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        for (long long j = lo; j < own_end; j += s_qs[__ldg(a.text + j + m)]) {  // ssabs.c:46-56
            if (s_x[m - 1] == __ldg(a.text + j + m - 1) && s_x[0] == __ldg(a.text + j)) {
                int i = m - 2;
                while (i > 0 && s_x[i] == __ldg(a.text + j + i)) i--;
                cnt += i <= 0;
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Fast-Search: table = bc[256] (all m characters) then gs[0..m] ---- */
template <bool W1>  // fs-w1.c:65-101: the same tables and shifts, the scan simply starts on the first window's last character
__global__ void __launch_bounds__(kTableThreads) table_fs_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_bc[256];
    __shared__ int raw_gs[SB_XSIZE + 1];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_bc[i] = a.table[i];
    for (int i = threadIdx.x; i <= a.m; i += kTableThreads) raw_gs[i] = a.table[256 + i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_bc(raw_bc);
    const Smem<int> s_gs(raw_gs);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        if (!W1) {  // fs.c:56: the first window is compared directly
            int i = 0;
            while (i < m && s_x[i] == __ldg(a.text + lo + i)) i++;
            cnt += i == m;
        }
        long long s = lo + (W1 ? m - 1 : m);  // index of the window's last character (fs.c:57 / fs-w1.c:87)
        while (s - m + 1 < own_end) {  // fs.c:58-66
            int k;
            while (s - m + 1 < own_end && (k = s_bc[__ldg(a.text + s)]) != 0) s += k;
            if (s - m + 1 >= own_end) break;
            int j = 2;
            while (j <= m && s_x[m - j] == __ldg(a.text + s - j + 1)) j++;
            cnt += j > m;
            s += s_gs[m - j + 1];
        }
    }
    table_reduce(a, cnt);
}

/* ---- GRASPm: table = bc[256], head[256], next[0..m] ---- */
__global__ void __launch_bounds__(kTableThreads) table_graspm_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_bc[256];
    __shared__ int raw_head[256];
    __shared__ int raw_next[SB_XSIZE + 1];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) {
        raw_bc[i] = a.table[i];
        raw_head[i] = a.table[256 + i];
    }
    for (int i = threadIdx.x; i <= a.m; i += kTableThreads) raw_next[i] = a.table[512 + i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_bc(raw_bc);
    const Smem<int> s_head(raw_head);
    const Smem<int> s_next(raw_next);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long j = lo + m - 1;
        while (j - (m - 1) < own_end) {  // graspm.c:52-67
            int k;
            while (j - (m - 1) < own_end && (k = s_bc[__ldg(a.text + j)]) != 0) j += k;
            if (j - (m - 1) >= own_end) break;
            for (int e = s_head[__ldg(a.text + j - 1)]; e >= 0; e = s_next[e]) {  // (j >= 1: m >= 1 and j >= lo + m - 1;
                const long long first = j - e;                                     //  for m == 1 the byte before local 0 is pad)
                if (first < lo || first >= own_end) continue;
                int i = 0;
                while (i < m && s_x[i] == __ldg(a.text + first + i)) i++;
                cnt += i == m;
            }
            j += m;
        }
    }
    table_reduce(a, cnt);
}

/* ---- Turbo-BM: table = bmBc[256] then bmGs[0..m) ---- */
__global__ void __launch_bounds__(kTableThreads) table_tbm_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_bc[256];
    __shared__ int raw_gs[SB_XSIZE];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_bc[i] = a.table[i];
    for (int i = threadIdx.x; i < a.m; i += kTableThreads) raw_gs[i] = a.table[256 + i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_bc(raw_bc);
    const Smem<int> s_gs(raw_gs);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        int u = 0, shift = m;
        for (long long j = lo; j < own_end; j += shift) {  // tbm.c:80-108
            int i = m - 1;
            unsigned c = 0;
            while (i >= 0 && s_x[i] == (c = __ldg(a.text + i + j))) {
                --i;
                if (u != 0 && i == m - 1 - shift) i -= u;
            }
            if (i < 0) {
                cnt++;
                shift = s_gs[0];
                u = m - shift;
            } else {
                const int v = m - 1 - i, turbo = u - v, bcs = s_bc[c] - m + 1 + i;
                shift = turbo > bcs ? turbo : bcs;
                if (s_gs[i] > shift) shift = s_gs[i];
                if (shift == s_gs[i]) {
                    u = (m - shift < v) ? m - shift : v;
                } else {
                    if (turbo < bcs && u + 1 > shift) shift = u + 1;
                    u = 0;
                }
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Apostolico-Crochemore: table = kmpNext[0..m]; a.q = ell ---- */
__global__ void __launch_bounds__(kTableThreads) table_ac_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_next[SB_XSIZE + 1];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i <= a.m; i += kTableThreads) raw_next[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_next(raw_next);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m, ell = a.q;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        int i = ell, k = 0;
        for (long long j = lo; j < own_end;) {  // ac.c:57-83
            while (i < m && s_x[i] == __ldg(a.text + i + j)) ++i;
            if (i >= m) {
                while (k < ell && s_x[k] == __ldg(a.text + j + k)) ++k;
                cnt += k >= ell;
            }
            j += i - s_next[i];
            if (i == ell) {
                k = k - 1 > 0 ? k - 1 : 0;
            } else if (s_next[i] <= ell) {
                k = s_next[i] > 0 ? s_next[i] : 0;
                i = ell;
            } else {
                k = ell;
                i = s_next[i];
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Zhu-Takaoka: table = bmGs[0..m) then ztBc[256][256] (read through L1/L2) ---- */
__global__ void __launch_bounds__(kTableThreads) table_zt_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_gs[SB_XSIZE];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < a.m; i += kTableThreads) raw_gs[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_gs(raw_gs);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const int32_t *zt = a.table + m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        for (long long j = lo; j < own_end;) {  // zt.c:88-99
            int i = m - 1;
            while (i >= 0 && s_x[i] == __ldg(a.text + i + j)) --i;
            if (i < 0) {
                cnt++;
                j += s_gs[0];
            } else {  // (m == 1: every ztBc entry is <= 1 = bmGs[0]; the reference reads y[j-1] for it)
                const int bc = m >= 2 ? __ldg(zt + __ldg(a.text + j + m - 2) * 256 + __ldg(a.text + j + m - 1)) : 1;
                const int gs = s_gs[i];
                j += gs > bc ? gs : bc;
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Skip Search on q-grams (skip3, skip5..8): table = next[0..m) then head[65536] (L1/L2); a.q = Q ---- */
__global__ void __launch_bounds__(kTableThreads) table_skipq_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_next[SB_XSIZE];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < a.m; i += kTableThreads) raw_next[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_next(raw_next);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m, Q = a.q, mq = m - Q + 1;
    const int32_t *head = a.table + m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        // skip5.c:50-57: every mq-th position is sampled; an occurrence at h is met by the one sample in [h, h+mq),
        // so the segment's own grid (from lo + mq - 1) finds every occurrence that starts in it
        for (long long j = lo + mq - 1; j - (m - Q) < own_end; j += mq) {
            unsigned h = 0;
            if (Q == 3) {
                h = ((unsigned)__ldg(a.text + j) << 4) + ((unsigned)__ldg(a.text + j + 1) << 2) + __ldg(a.text + j + 2);
            } else {
                for (int k = 0; k < Q; k++) h += (unsigned)__ldg(a.text + j + k) << (Q - 1 - k);
            }
            for (int e = __ldg(head + h); e >= 0; e = s_next[e]) {
                const long long st = j - e;
                if (st < lo || st >= own_end) continue;
                int k = 0;
                while (k < m && s_x[k] == __ldg(a.text + st + k)) k++;
                cnt += k == m;
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Weak Factor Recognition (wfr, wfrq2..8, wfr2..8 and the tuned twfr*): table = F as 65536 bits; a.q = characters per step ---- */
__global__ void __launch_bounds__(kTableThreads) table_wfr_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_F[2048];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 2048; i += kTableThreads) raw_F[i] = (unsigned)a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<unsigned> s_F(raw_F);
    const Smem<uint8_t> s_x(raw_x);
    const int plen = a.m, Q = a.q;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    auto factor = [&](unsigned h) { return (s_F[(h & 0xffffu) >> 5] >> (h & 31u)) & 1u; };
    auto whole = [&](long long st) {
        int k = 0;
        while (k < plen && s_x[k] == __ldg(a.text + st + k)) k++;
        return k == plen;
    };
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        if (Q <= 1) {  // wfr.c:73-92
            const int m = plen;
            long long j = lo + m - 1;
            while (j - (m - 1) < own_end) {
                unsigned h = __ldg(a.text + j);
                if (a.sh1) {  // twfr.c:57: while (!F[y[j]]) j += m - bounded by the segment instead of the sentinel
                    while (!factor(h) && j + m - (m - 1) < own_end) {
                        j += m;
                        h = __ldg(a.text + j);
                    }
                }
                const long long i = j - (m - 1);
                unsigned test;
                while ((test = factor(h)) && j > i) {
                    j--;
                    h = ((h << 2) + __ldg(a.text + j)) & 0xffffu;
                }
                if (j == i && test) cnt += whole(i);
                j += m;
            }
        } else {  // wfrq4.c:80-100: the window is the pattern cut to a multiple of Q
            const int m = plen - plen % Q;
            auto gram = [&](long long pos) {
                unsigned g = 0;
                for (int k = 0; k < Q; k++) g += (unsigned)__ldg(a.text + pos - k) << (2 * (Q - 1 - k));
                return g;
            };
            cnt += whole(lo);  // wfrq4.c:84: the first position is compared up front
            long long j = lo + m;
            while (j - m + 1 < own_end) {
                unsigned h = gram(j) & 0xffffu;
                if (a.sh1) {  // twfr4.c:63-66
                    while (!factor(h) && j + (m - Q + 1) - m + 1 < own_end) {
                        j += m - Q + 1;
                        h = gram(j) & 0xffffu;
                    }
                }
                const long long i = j - m + 1;
                unsigned test;
                while ((test = factor(h)) && j > i + Q - 1) {
                    j -= Q;
                    h = ((h << (2 * Q)) + gram(j)) & 0xffffu;
                }
                if (j == i + Q - 1 && test) cnt += whole(i);
                j += m - Q + 1;
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- bit-parallel window scans of the BNDM family: table = B[256] over the first k = min(m, 32) characters ---- */
/* the whole pattern at one start position (the comparison the reference makes up front and, for m > 32, after its
 * 32-character prefix automaton reports a window) */
__device__ __noinline__ unsigned whole_pattern(const uint8_t *text, const uint8_t *pat, long long st, int m)
{
    int i = 0;
    while (i < m && __ldg(pat + i) == __ldg(text + st + i)) i++;
    return i == m;
}

/* BNDM with q-grams (bndmq2.c:31-71, search_large :79-127; bndmq4.c, bndmq6.c): B as in bndm.c.  The window's last
 * Q characters are read at once (GRAMq); while factors survive, one more character to the left; a set top bit is a
 * pattern prefix - the whole window when the scan stands at its first character, else the next window starts there */
template <int Q>
__global__ void __launch_bounds__(kTableThreads) table_bndmq_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_B[256];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_B[i] = (unsigned)a.table[i];
    __syncthreads();
    const Smem<unsigned> s_B(raw_B);
    const int m = a.m, k = m < SB_WORD ? m : SB_WORD;
    const unsigned M = 1u << (k - 1);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        cnt += whole_pattern(a.text, a.pat, lo, m);  // bndmq2.c:57: position 0 is compared up front
        long long i = lo + k + 1 - Q;
        while (i - (k - Q) < own_end) {               // bndmq2.c:59-73
            unsigned D = s_B[__ldg(a.text + i)];
#pragma unroll
            for (int t = 1; t < Q; t++) D &= s_B[__ldg(a.text + i + t)] << t;
            if (D != 0) {
                long long j = i;
                const long long first = i - (k - Q);
                do {
                    if (D >= M) {
                        if (j > first) i = j - 1;
                        else if (first < own_end) cnt += m > SB_WORD ? whole_pattern(a.text, a.pat, first, m) : 1u;
                    }
                    j--;
                    D = (D << 1) & s_B[__ldg(a.text + j)];
                } while (D != 0);
            }
            i += k - Q + 1;
        }
    }
    table_reduce(a, cnt);
}

/* Forward NDM (fndm.c:28-60, search_large :68-104): the Shift-Or masks (bit j clear <=> x[j] == c, so.c:27-38)
 * drive a FORWARD factor scan from the window's last character; a clear top bit is a pattern suffix ending here and
 * the window is compared; when no factor survives the scan jumps m (k) ahead.  table = S[256], a.lim = ~0 << (k-1) */
__global__ void __launch_bounds__(kTableThreads) table_fndm_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_S[256];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_S[i] = (unsigned)a.table[i];
    __syncthreads();
    const Smem<unsigned> s_S(raw_S);
    const int m = a.m, k = m < SB_WORD ? m : SB_WORD;
    const unsigned top = a.lim;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long i = lo + k - 1;
        while (i - (k - 1) < own_end) {  // fndm.c:44-57
            unsigned D = s_S[__ldg(a.text + i)];
            while (D != ~0u && i - (k - 1) < own_end) {  // (windows that start at own_end or later are not this segment's)
                if (D < top) cnt += whole_pattern(a.text, a.pat, i - k + 1, m);
                i++;
                D = (D << 1) | s_S[__ldg(a.text + i)];
            }
            i += k;
        }
    }
    table_reduce(a, cnt);
}

/* Simplified BNDM (sbndm.c:28-84, search_large :92-146): B aligned to the top of the word, the window read right to
 * left until the factor dies or its first character is reached; after an occurrence the shift is the pattern's
 * period (sbndm.c:44-55 `restore[0]`).  table = B[256] (top aligned), a.sh1 = shift */
template <bool BMH>  // sbndm-bmh.c:27-84: + a Horspool-like skip on the character k behind the window while B[y[i]] == 0
__global__ void __launch_bounds__(kTableThreads) table_sbndm_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_B[512];
    for (int i = threadIdx.x; i < (BMH ? 512 : 256); i += kTableThreads) raw_B[i] = (unsigned)a.table[i];
    __syncthreads();
    const Smem<unsigned> s_B(raw_B);
    const int m = a.m, k = m < SB_WORD ? m : SB_WORD, shift = a.sh1;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        cnt += whole_pattern(a.text, a.pat, lo, m);  // sbndm.c:61
        long long i = lo + k;                         // the window's last character
        while (i - k + 1 < own_end) {                 // sbndm.c:63-83 (bounded by the segment, not by the sentinel)
            unsigned D = s_B[__ldg(a.text + i)];
            if (BMH) {                                // sbndm-bmh.c:61: hbc[c] = 2k - 1 - (last position of c), 2k if absent
                while (D == 0 && i - k + 1 < own_end) {
                    i += (int)s_B[256 + __ldg(a.text + i + k)];
                    D = s_B[__ldg(a.text + i)];
                }
                if (i - k + 1 >= own_end) break;
            } else if (m > SB_WORD && D == 0) {       // search_large's skip loop, sbndm.c:128
                i += k;
                continue;
            }
            long long j = i - 1;
            const long long first = i - k + 1;
            for (;;) {
                D = (D << 1) & s_B[__ldg(a.text + j)];
                if (!((j - first) && D)) break;
                j--;
            }
            if (D != 0) {
                cnt += m > SB_WORD ? whole_pattern(a.text, a.pat, first, m) : 1u;
                i += shift;
            } else {
                i = j + k;
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Tailed-Substring (ts.c:25-59): no table, the pattern in shared memory ---- */
/* phase 1 probes pattern position i and looks for the position whose previous equal character is furthest away
 * (dim, k); phase 2 keeps probing position k and shifts by dim.  Every segment runs both phases from scratch, as
 * the reference does from the start of the text. */
__global__ void __launch_bounds__(kTableThreads) table_ts_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    auto whole = [&](long long st) {
        int j = 0;
        while (j < m && s_x[j] == __ldg(a.text + st + j)) j++;
        return j == m;
    };
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long s = lo;
        int i = m - 1, k = m - 1, dim = 1;
        while (s < own_end && i - dim >= 0) {  // ts.c:33-47
            if (s_x[i] != __ldg(a.text + s + i)) {
                s++;
            } else {
                cnt += whole(s);
                int h = i - 1;
                const unsigned xi = s_x[i];
                while (h >= 0 && s_x[h] != xi) h--;
                if (dim < i - h) {
                    k = i;
                    dim = i - h;
                }
                s += i - h;
                i--;
            }
        }
        const unsigned xk = s_x[k];
        while (s < own_end) {                  // ts.c:49-57
            if (xk != __ldg(a.text + s + k)) {
                s++;
            } else {
                cnt += whole(s);
                s += dim;
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- KMP-Skip-Search (kmpskip.c:59-146): table = mpNext[0..m], kmpNext[0..m], z[256], list[0..m) ---- */
/* staged as 16-bit entries (every value lies in [-1, m]) */
__global__ void __launch_bounds__(kTableThreads) table_kmpskip_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ short raw_mp[SB_XSIZE + 1];
    __shared__ short raw_kmp[SB_XSIZE + 1];
    __shared__ short raw_list[SB_XSIZE];
    __shared__ short raw_z[256];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    const int m = a.m;
    for (int i = threadIdx.x; i <= m; i += kTableThreads) {
        raw_mp[i] = (short)a.table[i];
        raw_kmp[i] = (short)a.table[m + 1 + i];
    }
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_z[i] = (short)a.table[2 * (m + 1) + i];
    for (int i = threadIdx.x; i < m; i += kTableThreads) raw_list[i] = (short)a.table[2 * (m + 1) + 256 + i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<short> s_mp(raw_mp), s_kmp(raw_kmp), s_list(raw_list), s_z(raw_z);
    const Smem<uint8_t> s_x(raw_x);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        const long long nn = own_end + m - 1;  // text bytes the segment's windows may touch: [lo, nn)
        const int per = m - s_kmp[m];
        long long j = lo - 1, start, wall = lo, kmpStart;
        int i = -1, k;
        // kmpskip.c:83-90 / :104-111 / :121-128: the next sampled character whose bucket is not empty
        auto next_bucket = [&]() {
            do {
                j += m;
            } while (j < nn && s_z[__ldg(a.text + j)] < 0);
            if (j >= nn) return false;
            i = s_z[__ldg(a.text + j)];
            return true;
        };
        bool more = next_bucket();
        start = j - i;
        while (more && start < own_end) {  // kmpskip.c:92-143
            if (start > wall) wall = start;
            k = (int)(wall - start);
            while (k < m && s_x[k] == __ldg(a.text + k + start)) ++k;
            wall = start + k;
            if (k == m) {
                cnt += start >= lo;  // (a candidate left of the segment belongs to the previous thread)
                i -= per;
            } else {
                i = s_list[i];
            }
            if (i < 0 && !(more = next_bucket())) break;
            kmpStart = start + k - s_kmp[k];
            k = s_kmp[k];
            start = j - i;
            while (start < kmpStart || (kmpStart < start && start < wall)) {
                if (start < kmpStart) {
                    i = s_list[i];
                    if (i < 0 && !(more = next_bucket())) break;
                    start = j - i;
                } else {
                    kmpStart += k - s_mp[k];
                    k = s_mp[k];
                }
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Improved Occurrence Matcher (iom.c:28-56): table = ibc1[256], ibc2[256]; a.q = step1, a.sh1 = step2 ---- */
/* the shift is read from the character `step` positions ahead - up to m-1 behind the window; past the text that is
 * the zero pad, which can only concern windows that do not lie inside the text */
__global__ void __launch_bounds__(kTableThreads) table_iom_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_ibc[512];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 512; i += kTableThreads) raw_ibc[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_ibc(raw_ibc);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m, step1 = a.q, step2 = a.sh1;
    const unsigned last_ch = s_x[m - 1];
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        for (long long s = lo; s < own_end;) {  // iom.c:44-53
            if (last_ch == __ldg(a.text + s + m - 1)) {
                int i = 0;
                while (i < m && s_x[i] == __ldg(a.text + s + i)) i++;
                cnt += i == m;
                s += s_ibc[__ldg(a.text + s + step1)];
            } else {
                s += s_ibc[256 + __ldg(a.text + s + step2)];
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Quantum-Leap Quick-Search (qlqs.c:38-62): table = qsBc of x, then qsBc of the reversed x ---- */
__global__ void __launch_bounds__(kTableThreads) table_qlqs_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_qs[512];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 512; i += kTableThreads) raw_qs[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_qs(raw_qs);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m, z = m + m / 2;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        for (long long i = lo; i < own_end;) {  // qlqs.c:52-57
            int k = 0;
            while (k < m && s_x[k] == __ldg(a.text + i + k)) k++;
            cnt += k == m;
            const int f = s_qs[__ldg(a.text + i + m)], b = s_qs[256 + __ldg(a.text + i + z - 1)];
            i += f > z - b ? z : f;
        }
    }
    table_reduce(a, cnt);
}

/* ---- two-ended Berry-Ravindran scans: table = brBc of x, then brBc of the reversed x (2 x 256 KiB, L1/L2) ---- */
/* tsw.c:40-71 and tvsbs-w2.c:36-71 move one window up from the start and one down from the end until they cross;
 * a segment does that between its own first and last start position.  tvsbs-w2 looks at the first and last
 * character of both windows before comparing (:57-58). */
template <bool TVSBS>
__global__ void __launch_bounds__(kTableThreads) table_br2_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const int32_t *L = a.table, *R = a.table + 65536;
    const unsigned first_ch = s_x[0], last_ch = s_x[m - 1];
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    auto whole = [&](long long st) {
        int i = 0;
        while (i < m && s_x[i] == __ldg(a.text + st + i)) i++;
        return i == m;
    };
    if (lo < a.hi) {
        long long j = lo, e = min(lo + a.seg, a.hi) - 1;  // the first and the last start position of the segment
        while (j <= e) {
            bool look = true;
            if (TVSBS)
                look = (first_ch == __ldg(a.text + j) || first_ch == __ldg(a.text + e)) &&
                       (last_ch == __ldg(a.text + j + m - 1) || last_ch == __ldg(a.text + e + m - 1));
            if (look) {
                cnt += whole(j);
                if (j < e) cnt += whole(e);
            }
            j += __ldg(L + __ldg(a.text + j + m) * 256 + __ldg(a.text + j + m + 1));
            e -= __ldg(R + __ldg(a.text + e - 1) * 256 + __ldg(a.text + e - 2));
        }
    }
    table_reduce(a, cnt);
}

/* BXS (bxs.c:30-75, bxs2.c .. bxs8.c): BNDM over every (k-Q+1)-th q-gram; a set top bit away from the window start
 * moves the sample grid there and re-enters the bit (`(D<<1)|1`), at the window start the whole pattern is compared */
template <int Q>
__global__ void __launch_bounds__(kTableThreads) table_bxs_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_B[256];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_B[i] = (unsigned)a.table[i];
    __syncthreads();
    const Smem<unsigned> s_B(raw_B);
    const int len = a.m, k = len < SB_WORD ? len : SB_WORD, mq1 = k - Q + 1;
    const unsigned mask = 1u << (k - 1);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        for (long long i = lo + mq1 - 1; i - mq1 + 1 < own_end; i += mq1) {  // bxs2.c:53-72; the window starts at i-mq1+1
            unsigned D = s_B[__ldg(a.text + i + Q - 1)];
#pragma unroll
            for (int t = Q - 2; t >= 0; t--) D = (D << 1) & s_B[__ldg(a.text + i + t)];
            if (D) {
                long long j = i;
                const long long first = i - mq1;
                do {
                    j--;
                    if (D >= mask) {
                        if (j - first) i = j;
                        else if (first + 1 < own_end) cnt += whole_pattern(a.text, a.pat, first + 1, len);
                        D = ((D << 1) | 1u) & s_B[__ldg(a.text + j)];
                    } else {
                        D = (D << 1) & s_B[__ldg(a.text + j)];
                    }
                } while (D && j > first);
            }
        }
    }
    table_reduce(a, cnt);
}

/* Long BNDM (lbndm.c:42-85): the pattern is cut into groups of k = ceil(m/32) characters that share one bit, the
 * window is sampled every k-th character; a surviving top bit near the window start has the k alignments compared */
__global__ void __launch_bounds__(kTableThreads) table_lbndm_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_B[256];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_B[i] = (unsigned)a.table[i];
    __syncthreads();
    const Smem<unsigned> s_B(raw_B);
    const int m = a.m, k = (m - 1) / SB_WORD + 1, m1 = m - 1, m2 = m1 - k, rmd = m - (m / k) * k;
    const unsigned M = 1u << (SB_WORD - 1);
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        for (long long j = lo; j < own_end;) {  // lbndm.c:64-82
            unsigned D = s_B[__ldg(a.text + j + m1)];
            int last = (D & M) ? m - k - rmd : m - rmd;
            int l = m2;
            while (D) {
                D = (D << 1) & s_B[__ldg(a.text + j + l)];
                if (D & M) {
                    if (l < k + rmd) {  // verify (lbndm.c:29-40): the k alignments from j on, inside the segment
                        for (int t = 0; t < k && j + t < own_end; t++) cnt += whole_pattern(a.text, a.pat, j + t, m);
                        break;
                    }
                    last = l - (k + rmd) + 1;
                }
                l -= k;
            }
            j += last;
        }
    }
    table_reduce(a, cnt);
}

/* ---- Average-Optimal Shift-Or (aoso2.c:42-77, search_large :99-132; aoso4.c, aoso6.c): table = B[256];
 *      a.q = q, a.lim = mm.  q interleaved Shift-Or automata over every q-th pattern character share one word; the
 *      text is sampled every q-th character, a zero among the automata's top bits names the alignments to compare.
 *      Any sample grid works (an occurrence is met by the sub-pattern that falls on it), so a segment starts its own. ---- */
__global__ void __launch_bounds__(kTableThreads) table_aoso_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_B[256];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_B[i] = (unsigned)a.table[i];
    __syncthreads();
    const Smem<unsigned> s_B(raw_B);
    const int m = a.m, q = a.q, k = m < SB_WORD ? m : SB_WORD, mq = k / q;
    const unsigned mm = a.lim;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        const long long end = own_end + m - 1;  // text bytes of this segment's windows
        unsigned D = ~0u;
        for (long long j = lo; j < end; j += q) {  // aoso2.c:68-74
            D = ((D & ~mm) << 1) | s_B[__ldg(a.text + j)];
            if ((D & mm) != mm) {                  // verify, aoso2.c:28-40
                unsigned E = (D & mm) ^ mm;
                while (E) {
                    const int s = 31 - __clz(E);
                    const long long i = j - (long long)(mq - 1) * q - s / mq;
                    if (i >= lo && i < own_end) cnt += whole_pattern(a.text, a.pat, i, m);
                    E &= ~(1u << s);
                }
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Forward BOM (fbom.c:28-96, sfbom.c): table = the oracle of bom ((m+1) x 256 x 16 bit, m+1 terminal bytes,
 *      padding), then FT[256][256] at a.lim shorts: FT[b][a] for the character b BEHIND the window and its last
 *      character a = the state after reading b, a (a window one further right), or a shift coded as m + shift when
 *      "ab" is no factor; a = x[m-1] always enters the oracle with the window's own last character (fbom.c:53-64) ---- */
__global__ void __launch_bounds__(kTableThreads) table_fbom_kernel(const __grid_constant__ TableArgs a)
{
    const int m = a.m;
    const short *trans = reinterpret_cast<const short *>(a.table);
    const short *FT = trans + a.lim;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        cnt += whole_pattern(a.text, a.pat, lo, m);  // fbom.c:68
        long long j = lo + m;                         // the window's last character
        while (j - m + 1 < own_end) {                 // fbom.c:71-80
            int p;
            while ((p = __ldg(FT + __ldg(a.text + j + 1) * 256 + __ldg(a.text + j))) > m && j - m + 1 < own_end) j += p - m;
            if (p > m) break;
            long long i = j - 1;
            while (i >= j - m && (p = __ldg(trans + p * 256 + __ldg(a.text + i))) >= 0) i--;  // (state 0 has no exits)
            if (i < j - (m - 1)) {
                cnt += j - m + 1 < own_end;
                i++;
            }
            j = i + m;
        }
    }
    table_reduce(a, cnt);
}

/* ---- Shift Vector Matching (svm0.c:36-72, search_large :80-124): table = cv[256] over the first k = min(m, 31)
 *      characters (bit b clear <=> x[k-1-b] == c).  sv collects, for the next 32 window positions, which are already
 *      ruled out; a window is read right to left while position 0 is still possible, then the scan moves to the next
 *      possible position (bsf).  A segment starts with an empty vector. ---- */
__global__ void __launch_bounds__(kTableThreads) table_svm_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned raw_cv[256];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_cv[i] = (unsigned)a.table[i];
    __syncthreads();
    const Smem<unsigned> s_cv(raw_cv);
    const int m = a.m, k = m < SB_WORD - 1 ? m : SB_WORD - 1;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        cnt += whole_pattern(a.text, a.pat, lo, m);  // svm0.c:56
        unsigned sv = 0;
        for (long long s = lo + k; s - k + 1 < own_end;) {  // s: the window's last character (svm0.c:58-70)
            sv |= s_cv[__ldg(a.text + s)];
            int j = 1;
            while ((sv & 1u) == 0) {
                if (j >= k) {
                    cnt += m > k ? whole_pattern(a.text, a.pat, s - k + 1, m) : 1u;
                    break;
                }
                sv |= s_cv[__ldg(a.text + s - j)] >> j;
                j++;
            }
            const int adv = __ffs(~(sv >> 1));  // bsf + 1: bit 31 of ~(sv >> 1) is always set
            sv = adv < 32 ? sv >> adv : 0u;
            s += adv;
        }
    }
    table_reduce(a, cnt);
}

/* ---- Worst-Character rule (wc.c:75-96) and Worst-Occurrence Matcher (wom.c:58-80): table = the shift table for
 *      the position `pos` = a.q that FindBadPos / FindWorstOccurrence picks; compare, then shift on y[s + pos] ---- */
__global__ void __launch_bounds__(kTableThreads) table_wc_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_bc[256];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_bc[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_bc(raw_bc);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m, pos = a.q;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        for (long long s = lo; s < own_end;) {  // wc.c:84-93
            int i = 0;
            while (i < m && s_x[i] == __ldg(a.text + s + i)) i++;
            cnt += i == m;
            s += s_bc[__ldg(a.text + s + pos)];
        }
    }
    table_reduce(a, cnt);
}

/* ---- Probabilistic BMH (pbmh.c:29-61): table = hbc[256] (hor.c:26-30), then the comparison order v[0..m): the
 *      pattern positions sorted by the frequency of their character among the text's first 100 characters ---- */
__global__ void __launch_bounds__(kTableThreads) table_pbmh_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_hbc[256];
    __shared__ short raw_v[SB_XSIZE];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) raw_hbc[i] = a.table[i];
    for (int i = threadIdx.x; i < a.m; i += kTableThreads) raw_v[i] = (short)a.table[256 + i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_hbc(raw_hbc);
    const Smem<short> s_v(raw_v);
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        for (long long s = lo; s < own_end;) {  // pbmh.c:53-58
            int i = 0;
            while (i < m) {
                const int v = s_v[i];
                if (s_x[v] != __ldg(a.text + s + v)) break;
                i++;
            }
            cnt += i == m;
            s += s_hbc[__ldg(a.text + s + m - 1)];
        }
    }
    table_reduce(a, cnt);
}

/* ---- Backward SNR DAWG Matching (bsdm.c:57-76, bsdm4.c:68-93): table = pos[65536] (L1/L2); a.q = Q,
 *      a.sh1 = st, a.lim = len (the longest run of pattern q-grams without a repeat) ---- */
__global__ void __launch_bounds__(kTableThreads) table_bsdm_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<uint8_t> s_x(raw_x);
    const int m = a.m, Q = a.q, st = a.sh1, len = (int)a.lim, offset = st + len - 1, step = Q == 4 ? 2 : 1;
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    auto gram = [&](long long j) {
        unsigned h = 0;
        for (int t = 0; t < Q; t++) h += (unsigned)__ldg(a.text + j + t) << (step * (Q - 1 - t));
        return h & 0xffffu;
    };
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        long long j = lo + (Q == 1 ? offset : len - 1);  // bsdm.c:62 / bsdm4.c:73
        while (j - offset < own_end) {
            int i;
            while ((i = __ldg(a.table + gram(j))) < 0 && j - offset < own_end) j += len;  // the fast loop, bounded by the segment
            if (i < 0) break;
            int k = 1;
            while (k <= i && s_x[st + i - k] == __ldg(a.text + j - k)) k++;
            if (k > i) {  // the run's prefix up to this gram stands in the text
                if (k == len) {
                    const long long s = j - offset;
                    if (s >= lo && s < own_end) {
                        int t = 0;
                        while (t < m && s_x[t] == __ldg(a.text + s + t)) t++;
                        cnt += t == m;
                    }
                } else {
                    j -= k;
                }
            }
            j += len;
        }
    }
    table_reduce(a, cnt);
}

/* ---- Q-gram filtering: table = B[1 << (Q*S)]; a.q = Q, a.sh1 = S ---- */
constexpr int kQfSharedEntries = 8192;
__global__ void __launch_bounds__(kTableThreads) table_qf_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int raw_B[kQfSharedEntries];
    __shared__ uint8_t raw_x[SB_XSIZE + 8];
    const int Q = a.q, S = a.sh1, asize = 1 << (Q * S);
    const bool small = asize <= kQfSharedEntries;
    if (small)
        for (int i = threadIdx.x; i < asize; i += kTableThreads) raw_B[i] = a.table[i];
    stage_pattern_smem(raw_x, a);
    __syncthreads();
    const Smem<int> s_B(raw_B);
    const Smem<uint8_t> s_x(raw_x);
    const unsigned mask = (unsigned)asize - 1u;
    const int m = a.m, mq1 = m - Q + 1;
    const long long n = a.hi + m - 1;  // text bytes of this shard's own windows
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    auto gram = [&](long long i) {  // qf23.c:62 / qf33.c:58-60: ((..y[i+Q-1] << S) + ..) << S) + y[i]
        unsigned ch = __ldg(a.text + i + Q - 1);
        for (int t = Q - 2; t >= 0; t--) ch = (ch << S) + __ldg(a.text + i + t);
        return ch & mask;
    };
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        // qf23.c:61: the scan state is the index i alone, so a segment starts like the text does (i = mq1-1, at lo)
        for (long long i = lo + mq1 - 1; i <= n - Q && i - mq1 + 1 < own_end; i += mq1) {
            unsigned D = small ? (unsigned)s_B[gram(i)] : (unsigned)__ldg(a.table + gram(i));
            if (!D) continue;
            const long long j = i - mq1 + Q;
            for (;;) {                                      // qf23.c:65-72 ("more:")
                i -= Q;
                if (i >= j) {
                    D &= small ? (unsigned)s_B[gram(i)] : (unsigned)__ldg(a.table + gram(i));
                    if (D == 0) break;
                } else {                                    // qf23.c:73-81: verify the surviving alignments
                    i = j;
                    long long k = j - Q + 1, top = j > n - m ? n - m : j;
                    if (k < lo) k = lo;                     // (start positions of this segment only)
                    if (top > own_end - 1) top = own_end - 1;
                    for (; k <= top; k++) {
                        int t = 0;
                        while (t < m && s_x[t] == __ldg(a.text + k + t)) t++;
                        if (t == m) cnt++;
                    }
                    break;
                }
            }
        }
    }
    table_reduce(a, cnt);
}

}  // namespace sb
