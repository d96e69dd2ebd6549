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
 *
 * Parallelisation: the scan loops are inherently sequential, so the text is cut into segments of `seg` start
 * positions, one THREAD per segment, each reading m-1 bytes past its segment (the same m-1 halo rule as the shards).
 * A thread that starts its automaton at the segment's first byte finds exactly the occurrences that start inside
 * the segment.  Lanes of a warp therefore read addresses `seg` bytes apart: every load instruction touches 32
 * sectors and the kernels are bound by the L1/LSU sector rate, not by HBM - measured 0.2-1.3 TB/s against 7.2 TB/s
 * for the fingerprint kernels, which is why these are a selectable mode and not the default.
 * What they return is bit-identical to everything else: the count of bf.c:33-36.
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
    int seg;                      // start positions per thread (multiple of 4)
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

/* pattern bytes into shared memory (padded to a word) */
__device__ __forceinline__ void stage_pattern_smem(uint8_t *sx, const TableArgs &a)
{
    for (int i = threadIdx.x; i < a.m; i += kTableThreads) sx[i] = a.pat[i];
}

/* ---- Knuth-Morris-Pratt / Morris-Pratt: table = kmpNext[0..m] or mpNext[0..m] ---- */
__global__ void __launch_bounds__(kTableThreads) table_kmp_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int s_next[SB_XSIZE + 1];
    __shared__ uint8_t s_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i <= a.m; i += kTableThreads) s_next[i] = a.table[i];
    stage_pattern_smem(s_x, a);
    __syncthreads();
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);   // starts [lo, own_end)
        const long long end = own_end + a.m - 1;           // text bytes [lo, end)
        int i = 0;
        uint32_t w = 0;
        for (long long j = lo; j < end; j++) {
            if ((j & 3) == 0) w = __ldg(reinterpret_cast<const uint32_t *>(a.text + j));  // lo is a multiple of 4
            const int c = (int)(w & 0xffu);
            w >>= 8;
            while (i > -1 && (int)s_x[i] != c) i = s_next[i];  // kmp.c:57-58
            i++;
            if (i >= a.m) {                                     // kmp.c:61-64: OUTPUT(j - i) with j already advanced
                cnt++;                                          // the start j+1-m is >= lo (i began at 0) and < own_end
                i = s_next[i];
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- Horspool: table = hbc[256] ---- */
__global__ void __launch_bounds__(kTableThreads) table_hor_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int s_hbc[256];
    __shared__ uint8_t s_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) s_hbc[i] = a.table[i];
    stage_pattern_smem(s_x, a);
    __syncthreads();
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
template <bool AND>
__global__ void __launch_bounds__(kTableThreads) table_so_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ unsigned s_S[256];
    __shared__ uint8_t s_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) s_S[i] = (unsigned)a.table[i];
    stage_pattern_smem(s_x, a);
    __syncthreads();
    const long long lo = ((long long)blockIdx.x * kTableThreads + threadIdx.x) * a.seg;
    const int k = a.m < SB_WORD ? a.m : SB_WORD;  // characters in the automaton (search_large: the 32-char prefix)
    unsigned cnt = 0;
    if (lo < a.hi) {
        const long long own_end = min(lo + a.seg, a.hi);
        const long long end = own_end + k - 1;  // the prefix automaton reads [lo, end)
        unsigned D = AND ? 0u : ~0u;
        uint32_t w = 0;
        for (long long j = lo; j < end; j++) {
            if ((j & 3) == 0) w = __ldg(reinterpret_cast<const uint32_t *>(a.text + j));
            const unsigned c = w & 0xffu;
            w >>= 8;
            bool hit;
            if (AND) {
                D = ((D << 1) | 1u) & s_S[c];  // sa.c:50
                hit = (D & a.lim) != 0;        // sa.c:51 (F = 1 << (k-1))
            } else {
                D = (D << 1) | s_S[c];         // so.c:55
                hit = D < a.lim;               // so.c:56
            }
            if (hit) {
                const long long st = j - k + 1;  // >= lo: the automaton started empty at lo
                int i = k;                       // so.c:88-91: the bytes after the prefix
                while (i < a.m && s_x[i] == __ldg(a.text + st + i)) i++;
                if (i == a.m) cnt++;
            }
        }
    }
    table_reduce(a, cnt);
}

/* ---- HASHq: table = shift[256]; a.q = 3, 5 or 8; a.sh1 ---- */
__global__ void __launch_bounds__(kTableThreads) table_hashq_kernel(const __grid_constant__ TableArgs a)
{
    __shared__ int s_shift[256];
    __shared__ uint8_t s_x[SB_XSIZE + 8];
    for (int i = threadIdx.x; i < 256; i += kTableThreads) s_shift[i] = a.table[i];
    stage_pattern_smem(s_x, a);
    __syncthreads();
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

}  // namespace sb
