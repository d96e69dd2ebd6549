/*
 * Hand-written sm_100a scan kernels for the exact-string-matching hot path.
 *
 * What every kernel computes is the result of src/algos/bf.c:33-36 (reference tree): the
 * number of start positions j in the shard's owned range with x[0..m) == y[j..j+m), and,
 * when asked, those positions.  How they get there is B200-first:
 *
 *   - the text is streamed from HBM exactly once with 128-bit coalesced, L1-bypassing loads
 *     (LDG.E.128.NA), 4 independent loads in flight per thread, one warp per 2 KiB group,
 *     groups strided across a persistent grid;
 *   - a packed candidate filter runs entirely in registers:
 *       * scan_shift  (m <= 14): the first min(m,8) pattern bytes are compared exactly at all
 *         16 start positions of a vector using funnel-shifted words (SHF + LOP3), the 8-byte
 *         halo comes from the neighbouring lane by shuffle, not from memory;
 *       * scan_block<B> (m >= 2B-1): every occurrence covers one B-aligned text block
 *         completely, so only ALIGNED blocks are fingerprinted (IMAD on the FMA pipe) and
 *         compared with the B fingerprints of the pattern's B-grams at offsets a..a+B-1
 *         (one ISETP.EQ.OR per start position on the ALU pipe, no shifts, no halo);
 *   - candidates (rare) go to a warp-cooperative exact verification that reads the text
 *     through L2 with unaligned 32-bit windows, 128 bytes per step;
 *   - counts are reduced with warp shuffles and ONE atomicAdd per block.
 *
 * Nothing here is a dense contraction: tensor cores / TMEM are deliberately unused; the bound
 * is HBM bandwidth, with integer issue rate as the secondary limit.
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sb {

constexpr int kThreads = 256;            // threads per CTA
constexpr int kUnroll = 4;               // 128-bit loads in flight per thread
constexpr int kGroupBytes = 32 * 16 * kUnroll;  // bytes one warp consumes per iteration (2 KiB)
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
    const uint4 *text;            // shard buffer, 16-B aligned, zero-padded to a group multiple + 1 group
    const uint32_t *pat;          // pattern as little-endian words, zero padded
    unsigned long long *counters; // [0] match count, [1] position cursor, [2] verified candidates
    unsigned long long *pos;      // position sink (unsorted) or nullptr
    unsigned long long pos_cap;
    unsigned long long base;      // global offset of local byte 0
    long long hi;                 // owned local start positions are [0, hi)
    unsigned long long ngroups;   // 2 KiB groups to scan
    int m;
    int anchor;                   // byte offset a of the fingerprinted sub-pattern
    int nwords;                   // ceil(m/4)
    uint32_t tailmask;            // mask for the last pattern word
    uint32_t H[32];               // block kernel: fingerprints of x[a+d .. a+d+B), d < B
    uint32_t X0, X1, M0, M1;      // shift kernel: first 8 pattern bytes (pre-masked) and their masks
};

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

/* 32-bit little-endian window at an arbitrary byte offset (two aligned loads + funnel shift) */
__device__ __forceinline__ uint32_t text_window32(const uint4 *text, long long off)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(text) + (off >> 2);
    uint32_t a = __ldg(w), b = __ldg(w + 1);
    return __funnelshift_r(a, b, (unsigned)(off & 3) * 8u);
}

/* Exact comparison of x[0..m) with text[j..j+m), one 32-bit window per lane per step.
 * All 32 lanes must call it with the same j. */
__device__ __forceinline__ bool warp_verify(const ScanArgs &a, long long j, unsigned lane)
{
    const int nwords = a.nwords;
    for (int base = 0; base < nwords; base += 32) {
        const int k = base + (int)lane;
        bool ok = true;
        if (k < nwords) {
            uint32_t diff = text_window32(a.text, j + 4ll * k) ^ __ldg(a.pat + k);
            if (k == nwords - 1) diff &= a.tailmask;
            ok = (diff == 0);
        }
        if (!__all_sync(kFull, ok)) return false;
    }
    return true;
}

__device__ __forceinline__ void emit_match(const ScanArgs &a, long long j)
{
    if (a.pos) {
        unsigned long long idx = atomicAdd(a.counters + 1, 1ull);
        if (idx < a.pos_cap) a.pos[idx] = a.base + (unsigned long long)j;
    }
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

/* ===================================================================== scan_block<B> */

/* Slow path of scan_block: lane `src` of this warp saw a fingerprint hit in its vector.
 * The whole warp re-derives which (block, d) slots hit and verifies each candidate. */
template <int B>
__device__ __noinline__ void block_candidates(const ScanArgs &a, uint4 v, uint32_t hpair, long long voff,
                                              bool hit, unsigned lane, unsigned &cnt, unsigned &nver)
{
    unsigned pending = __ballot_sync(kFull, hit);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const uint32_t sx = __shfl_sync(kFull, v.x, src), sy = __shfl_sync(kFull, v.y, src);
        const uint32_t sz = __shfl_sync(kFull, v.z, src), sw = __shfl_sync(kFull, v.w, src);
        const uint32_t sh = __shfl_sync(kFull, hpair, src);
        const long long soff = __shfl_sync(kFull, voff, src);
        // 16 (block, d) slots per vector; lanes 0..15 each test one
        const int slot = (int)(lane & 15);
        uint32_t h;
        int d;
        long long t;
        if (B == 4) {
            const int blk = slot >> 2;
            d = slot & 3;
            h = blk == 0 ? sx : blk == 1 ? sy : blk == 2 ? sz : sw;
            t = soff + 4 * blk;
        } else if (B == 8) {
            const int blk = slot >> 3;
            d = slot & 7;
            h = blk == 0 ? fp8(sx, sy) : fp8(sz, sw);
            t = soff + 8 * blk;
        } else if (B == 16) {
            d = slot;
            h = fp16(sx, sy, sz, sw);
            t = soff;
        } else {
            d = slot + 16 * (src & 1);
            h = sh;
            t = soff & ~31ll;
        }
        unsigned slots = __ballot_sync(kFull, lane < 16 && h == a.H[d]);
        while (slots) {
            const int s = __ffs(slots) - 1;
            slots &= slots - 1;
            const long long ts = __shfl_sync(kFull, t, s);
            const int ds = __shfl_sync(kFull, d, s);
            const long long j = ts - ds - a.anchor;
            if (j >= 0 && j < a.hi) {
                nver += (lane == 0);
                if (warp_verify(a, j, lane) && lane == 0) {
                    cnt++;
                    emit_match(a, j);
                }
            }
        }
    }
}

template <int B>
__global__ void __launch_bounds__(kThreads) scan_block_kernel(const __grid_constant__ ScanArgs a)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * kThreads) >> 5;
    unsigned cnt = 0, nver = 0;

    // the 16 fingerprints this lane compares against (B==32: even lanes d<16, odd lanes d>=16)
    uint32_t H[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (B == 32)
            H[i] = (lane & 1) ? a.H[16 + i] : a.H[i];
        else
            H[i] = a.H[i % B];
    }

    for (unsigned long long g = warp; g < a.ngroups; g += nwarps) {
        const uint4 *p = a.text + g * (kGroupBytes / 16) + lane;
        uint4 v[kUnroll];
#pragma unroll
        for (int c = 0; c < kUnroll; c++) v[c] = ldg_stream(p + c * 32);
#pragma unroll
        for (int c = 0; c < kUnroll; c++) {
            bool hit = false;
            uint32_t hp = 0;
            if (B == 4) {
#pragma unroll
                for (int d = 0; d < 4; d++)
                    hit |= (v[c].x == H[d]) | (v[c].y == H[d]) | (v[c].z == H[d]) | (v[c].w == H[d]);
            } else if (B == 8) {
                const uint32_t h0 = fp8(v[c].x, v[c].y), h1 = fp8(v[c].z, v[c].w);
#pragma unroll
                for (int d = 0; d < 8; d++) hit |= (h0 == H[d]) | (h1 == H[d]);
            } else if (B == 16) {
                const uint32_t h = fp16(v[c].x, v[c].y, v[c].z, v[c].w);
#pragma unroll
                for (int d = 0; d < 16; d++) hit |= (h == H[d]);
            } else {
                const uint32_t mine = fp16(v[c].x, v[c].y, v[c].z, v[c].w);
                const uint32_t other = __shfl_xor_sync(kFull, mine, 1);
                hp = (lane & 1) ? fp32(other, mine) : fp32(mine, other);
#pragma unroll
                for (int d = 0; d < 16; d++) hit |= (hp == H[d]);
            }
            if (__any_sync(kFull, hit)) {
                const long long voff = (long long)(g * kGroupBytes) + c * 512 + lane * 16;
                block_candidates<B>(a, v[c], hp, voff, hit, lane, cnt, nver);
            }
        }
    }
    block_reduce_count(a, cnt, nver);
}

/* ===================================================================== scan_shift */

/* Per-lane tail of scan_shift when m > 8 or positions are wanted: `mm` has bit p set for every
 * start position p of this lane's vector whose first min(m,8) bytes match. */
__device__ __noinline__ void shift_candidates(const ScanArgs &a, unsigned mm, long long voff, unsigned &cnt,
                                              unsigned &nver)
{
    const unsigned char *tb = reinterpret_cast<const unsigned char *>(a.text);
    const unsigned char *pb = reinterpret_cast<const unsigned char *>(a.pat);
    while (mm) {
        const int p = __ffs(mm) - 1;
        mm &= mm - 1;
        const long long j = voff + p;
        if (j >= a.hi) break;
        bool ok = true;
        if (a.m > 8) {
            nver++;
            for (int k = 8; k < a.m; k++)
                if (__ldg(tb + j + k) != __ldg(pb + k)) {
                    ok = false;
                    break;
                }
        }
        if (ok) {
            cnt++;
            emit_match(a, j);
        }
    }
}

/* HI: compare a second word (m > 4).  MASK: collect a per-position mask and post-process
 * (needed for m > 8, for position output and for vectors straddling the owned range). */
template <bool HI>
__device__ __forceinline__ unsigned shift_match_mask(const ScanArgs &a, const uint32_t (&W)[6])
{
    unsigned mm = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t lo = r ? __funnelshift_r(W[q], W[q + 1], 8 * r) : W[q];
            uint32_t t = (lo & a.M0) ^ a.X0;
            if (HI) {
                const uint32_t hi = r ? __funnelshift_r(W[q + 1], W[q + 2], 8 * r) : W[q + 1];
                t |= (hi & a.M1) ^ a.X1;
            }
            mm |= (t == 0 ? 1u : 0u) << (4 * q + r);
        }
    }
    return mm;
}

template <bool HI>
__device__ __forceinline__ unsigned shift_match_count(const ScanArgs &a, const uint32_t (&W)[6])
{
    unsigned c = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t lo = r ? __funnelshift_r(W[q], W[q + 1], 8 * r) : W[q];
            uint32_t t = (lo & a.M0) ^ a.X0;
            if (HI) {
                const uint32_t hi = r ? __funnelshift_r(W[q + 1], W[q + 2], 8 * r) : W[q + 1];
                t |= (hi & a.M1) ^ a.X1;
            }
            c += (t == 0);
        }
    }
    return c;
}

template <bool HI, bool MASK>
__global__ void __launch_bounds__(kThreads) scan_shift_kernel(const __grid_constant__ ScanArgs a)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * kThreads) >> 5;
    unsigned cnt = 0, nver = 0;

    for (unsigned long long g = warp; g < a.ngroups; g += nwarps) {
        const uint4 *p = a.text + g * (kGroupBytes / 16) + lane;
        uint4 v[kUnroll];
#pragma unroll
        for (int c = 0; c < kUnroll; c++) v[c] = ldg_stream(p + c * 32);
        // halo for the last chunk's lane 31: the first 8 bytes of the next group (lane 0 fetches them)
        uint2 gn = make_uint2(0u, 0u);
        if (lane == 0) gn = __ldg(reinterpret_cast<const uint2 *>(a.text + (g + 1) * (kGroupBytes / 16)));
#pragma unroll
        for (int c = 0; c < kUnroll; c++) {
            // next 8 bytes after this lane's vector: lane+1's (x,y); lane 31 takes lane 0 of the next chunk
            uint32_t sx, sy;
            if (c + 1 < kUnroll) {
                sx = lane == 0 ? v[c + 1].x : v[c].x;
                sy = lane == 0 ? v[c + 1].y : v[c].y;
            } else {
                sx = lane == 0 ? gn.x : v[c].x;
                sy = lane == 0 ? gn.y : v[c].y;
            }
            const uint32_t nx = __shfl_sync(kFull, sx, (lane + 1) & 31);
            const uint32_t ny = __shfl_sync(kFull, sy, (lane + 1) & 31);
            const uint32_t W[6] = {v[c].x, v[c].y, v[c].z, v[c].w, nx, ny};
            const long long voff = (long long)(g * kGroupBytes) + c * 512 + lane * 16;
            if (!MASK && voff + 16 <= a.hi) {
                cnt += shift_match_count<HI>(a, W);
            } else if (voff < a.hi) {
                const unsigned mm = shift_match_mask<HI>(a, W);
                if (mm) shift_candidates(a, mm, voff, cnt, nver);
            }
        }
    }
    block_reduce_count(a, cnt, nver);
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

}  // namespace sb
