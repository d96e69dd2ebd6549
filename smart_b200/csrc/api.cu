/*
 * C-ABI implementation: device-resident (optionally sharded) text handles, the search entry
 * points, and the host-buffer streaming path.  See include/smart_b200.h for the contract and
 * the reference interfaces each entry point replaces.
 *
 * There is deliberately NO CPU fallback in this file: if CUDA is unavailable every entry point
 * that would scan returns SB_ERR_CUDA with a message in sb_last_error().
 */
#include "sb_internal.h"
#include "scan_kernels.cuh"
#include "table_kernels.cuh"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <vector>


namespace {

thread_local char g_err[512] = "";
thread_local sb_timing g_timing;

void set_err(const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(g_err, sizeof g_err, fmt, a, b);
}

#define CU_TRY(expr, ret)                                                     \
    do {                                                                      \
        cudaError_t e_ = (expr);                                              \
        if (e_ != cudaSuccess) {                                              \
            set_err("CUDA error in %s: %s", #expr, cudaGetErrorString(e_));   \
            return ret;                                                       \
        }                                                                     \
    } while (0)

double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

constexpr size_t kHalo = SB_XSIZE - 1;        // right halo kept with every shard
constexpr size_t kShardAlign = 4096;          // shard bases are multiples of this
constexpr int kMaxShards = 16;
constexpr size_t kPatBytes = 4352;            // device pattern buffer (>= SB_XSIZE + 8, multiple of 256)
#ifndef SB_U
#define SB_U 4
#endif
constexpr int kU = SB_U;                      // 256-bit loads in flight per thread (scan_block / scan_window8)
constexpr size_t kGroupLong = kU * 1024;      // bytes per warp iteration of those kernels
constexpr size_t kGroupPacked = 4 * 1024;     // ... of the kernels whose U is fixed at 4 (scan_swar, scan_packed*, scan_pswar)
constexpr int kFilterShort = 0, kFilterWindow8 = 1, kFilterWindow16 = 3;  // 2 = the 12-byte window
constexpr int kFilterPair = 20;               // scan_block<2,L>: windows at even positions, L = 4..8 in Plan::flen
/* scan_packed<8> moves ~26 T characters/s on a 2-bit copy and ~49 T on a 1-bit one (~13 T on a 4-bit one);
 * scan_skip passes that from m ~ 450 resp. ~1200 on (and beats the 4-bit copy as soon as it applies, m >= 167) (measured, profiles/packed_sweep_r1_*.jsonl) */
constexpr int kPackedBeatsSkipBelow4 = 167, kPackedBeatsSkipBelow2 = 448, kPackedBeatsSkipBelow1 = 1200;
constexpr int kFilterPacked = 300;             // + K (301, 302, 304, 308): scan_packed<K> on the 2-bit recoding
constexpr int kFilterSkip = 100;              // + gram bytes (108, 116, 132): scan_skip<G>
constexpr int kFilterPswar = 500;             // scan_pswar<BITS,KF>: reported as 500 + KF (1 bit) / 540 + KF (2 bits)
/* scan_pswar covers m <= 30 on a 1-bit copy and m <= 22 on a 2-bit one; beyond, the half-word bitmap kernels are faster */
constexpr int kPswarMaxM1 = 30, kPswarMaxM2 = 22;
constexpr size_t kAuxWords = sb::kBloomWords + 2 * sb::kSkipTableSize;
constexpr size_t kTableInts = (size_t)(SB_XSIZE + 1) * 128 + SB_XSIZE / 4 + 64 + 32768;  // (+ ebom's FT)  // SB_FLAG_TABLE_SCAN: the largest table is aut's (m+1) x 256 x 16 bit (qf: 2^18 ints)

constexpr int kTickets = 256;                 // outstanding asynchronous searches (sb_search64_submit)
constexpr int kSyncSlot = kTickets;           // pattern/counter slot of the blocking paths (positions, multi)
constexpr int kBatchJobs = 32;                // searches per pass of sb_search_host_batch
constexpr int kBatchSlot0 = kTickets + 1;     // their pattern/counter slots
constexpr int kSlots = kTickets + 1 + kBatchJobs;
constexpr int kAuxSlots = 40;                 // cached Bloom/table blocks per device (>= kBatchJobs: a batch pins one per job)

/* One cached aux block (Bloom bitmap + exact fingerprint table of a scan_skip / scan_packed plan).  The
 * driver's algorithm loop searches the same pattern with many algorithms: the block is built and uploaded once. */
struct AuxSlot {
    uint32_t *d = nullptr, *h = nullptr;      // device block, pinned staging
    cudaEvent_t staged = nullptr;             // fired once the H2D copy from h has run: h may be rewritten
    cudaEvent_t last_use = nullptr;           // fired once the last kernel reading d has finished
    bool used = false, valid = false;
    bool half = false;                        // scan_packed_half layout (16-bit fingerprints index the bitmap)
    unsigned long long stamp = 0;             // LRU clock
    std::vector<uint32_t> list;               // the fingerprints the block was built from (cache key)
};

struct DevCtx {                               // per-device scratch, created on first use
    bool ready = false;
    int dev = -1;
    int sms = 0;
    cudaStream_t stream = nullptr;            // scan kernels
    cudaStream_t ustream = nullptr;           // pattern / aux uploads of the asynchronous path (overlap the previous scan)
    cudaStream_t dstream = nullptr;           // counter read-backs of the asynchronous path
    cudaEvent_t e0 = nullptr, e1 = nullptr;   // blocking paths
    cudaEvent_t t0 = nullptr, t1 = nullptr;   // sb_stream_timer_*
    // slot s of kSlots: pattern words, 4 x u64 counters (device + pinned mirrors); slot kSyncSlot and the kBatchJobs
    // slots after it belong to the blocking paths.  The counters of the ticket slots are never zeroed: `prev` remembers what the slot
    // held after its previous search and the result is the difference (no memset node per search).
    uint32_t *d_pat = nullptr, *h_pat = nullptr;
    unsigned long long *d_counters = nullptr, *h_counters = nullptr;
    unsigned long long prev[kTickets][4] = {{0}};
    cudaEvent_t ev_up[kTickets] = {nullptr}, ev0[kTickets] = {nullptr}, ev1[kTickets] = {nullptr}, ev_done[kTickets] = {nullptr};
    AuxSlot aux[kAuxSlots];
    unsigned long long aux_clock = 0;
    unsigned long long *d_pos = nullptr;      // sorted position output, grown on demand
    size_t pos_cap = 0;
    uint32_t *d_posbits = nullptr;            // position bitmap of one window of start positions (kPosWindow bits)
    unsigned *d_blocksum = nullptr;           // per-CTA popcounts of the bitmap
    unsigned long long *d_blockoff = nullptr; // their exclusive scan (+ the running output offset)
    unsigned long long *d_written = nullptr;  // running output offset = occurrences seen so far on this device
    int occ_packed[4] = {0};                  // ... for scan_packed<K>, K = 1, 2, 4, 8
    int occ_packed_half[2] = {0};             // ... for scan_packed_half<KH>, KH = 1, 2
    int occ_pair[6] = {0};                    // ... for scan_block<2,L>, index L-3
    int occ_block[7] = {0};                   // resident CTAs/SM for scan_block<B,L>, index block_index()
    int occ_window[24] = {0};                 // ... for scan_window<NW,OFF,DENSE>, index window_index()
    int occ_swar[8] = {0};                    // ... for scan_swar<M, DENSE>, index (M-1)*2 + DENSE
    mutable int occ_pswar[2][33][2] = {{{0}}};  // ... for scan_pswar<BITS,KF,EXACT>, filled on first use
    unsigned long long *d_hist = nullptr;     // 256-bin byte histogram scratch
    int32_t *d_table = nullptr;               // SB_FLAG_TABLE_SCAN: the algorithm's table (kTableInts ints)
    uint32_t *d_mpat = nullptr;               // scan_multi: pattern set (words), grown on demand
    size_t mpat_words = 0;
    unsigned long long *d_mcounts = nullptr;  // scan_multi: per-pattern counts
    size_t mcounts_cap = 0;
    int occ_multi[3] = {0, 0, 0};             // resident CTAs/SM for scan_multi<4,8,16>
    int occ_skip[3] = {0, 0, 0};              // resident CTAs/SM for scan_skip<8,16,32>
    // host-buffer streaming path
    uint8_t *d_chunk[2] = {nullptr, nullptr};   // each preceded by kPadBytes of zeros (see text_alloc)
    size_t chunk_cap = 0;
    cudaStream_t cstream[2] = {nullptr, nullptr};

    uint32_t *pat_dev(int slot) const { return d_pat + (size_t)slot * (kPatBytes / 4); }
    uint32_t *pat_host(int slot) const { return h_pat + (size_t)slot * (kPatBytes / 4); }
    unsigned long long *cnt_dev(int slot) const { return d_counters + 4 * (size_t)slot; }
    unsigned long long *cnt_host(int slot) const { return h_counters + 4 * (size_t)slot; }
};

DevCtx g_ctx[64];
std::mutex g_mu;

/* scan_block<B,L> variants: disjoint blocks (L == B) 4,8,16,32 and overlapping windows <4,8> <4,16> <8,16> */
constexpr int block_index(int B, int L)
{
    return L == B ? (B == 4 ? 0 : B == 8 ? 1 : B == 16 ? 2 : 3) : (B == 4 && L == 8) ? 4 : (B == 4 && L == 16) ? 5 : 6;
}

constexpr int window_index(int nw, int off, bool dense) { return ((nw - 2) * 4 + (off - 1)) * 2 + (dense ? 1 : 0); }

template <typename K>
int occupancy(K kernel)
{
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, sb::kThreads, 0) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

DevCtx *get_ctx(int dev)
{
    if (dev < 0 || dev >= 64) {
        set_err("device index out of range%s%s");
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    DevCtx &c = g_ctx[dev];
    if (c.ready) {
        CU_TRY(cudaSetDevice(dev), nullptr);
        return &c;
    }
    CU_TRY(cudaSetDevice(dev), nullptr);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, dev), nullptr);
    if (prop.major < 10) {
        set_err("smart_b200 needs an sm_100a device (B200); found %s%s", prop.name);
        return nullptr;
    }
    c.dev = dev;
    c.sms = prop.multiProcessorCount;
    if (const char *e = getenv("SMART_B200_L2_FETCH"))  // tuning: DRAM->L2 fetch granularity hint (32/64/128)
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
    CU_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking), nullptr);
    CU_TRY(cudaStreamCreateWithFlags(&c.ustream, cudaStreamNonBlocking), nullptr);
    CU_TRY(cudaStreamCreateWithFlags(&c.dstream, cudaStreamNonBlocking), nullptr);
    CU_TRY(cudaStreamCreateWithFlags(&c.cstream[0], cudaStreamNonBlocking), nullptr);
    CU_TRY(cudaStreamCreateWithFlags(&c.cstream[1], cudaStreamNonBlocking), nullptr);
    CU_TRY(cudaEventCreate(&c.e0), nullptr);
    CU_TRY(cudaEventCreate(&c.e1), nullptr);
    CU_TRY(cudaEventCreate(&c.t0), nullptr);
    CU_TRY(cudaEventCreate(&c.t1), nullptr);
    CU_TRY(cudaMalloc(&c.d_pat, kSlots * kPatBytes), nullptr);
    CU_TRY(cudaHostAlloc(&c.h_pat, kSlots * kPatBytes, cudaHostAllocDefault), nullptr);
    CU_TRY(cudaMalloc(&c.d_counters, kSlots * 4 * sizeof(unsigned long long)), nullptr);
    CU_TRY(cudaHostAlloc(&c.h_counters, kSlots * 4 * sizeof(unsigned long long), cudaHostAllocDefault), nullptr);
    CU_TRY(cudaMemset(c.d_counters, 0, kSlots * 4 * sizeof(unsigned long long)), nullptr);
    memset(c.h_pat, 0, kSlots * kPatBytes);
    CU_TRY(cudaMalloc(&c.d_hist, 256 * sizeof(unsigned long long)), nullptr);
    c.occ_skip[0] = occupancy(sb::scan_skip_kernel<8>);
    c.occ_skip[1] = occupancy(sb::scan_skip_kernel<16>);
    c.occ_skip[2] = occupancy(sb::scan_skip_kernel<32>);
    c.occ_multi[0] = occupancy(sb::scan_multi_kernel<4, kU>);
    c.occ_multi[1] = occupancy(sb::scan_multi_kernel<8, kU>);
    c.occ_multi[2] = occupancy(sb::scan_multi_kernel<16, kU>);
    c.occ_packed[0] = occupancy(sb::scan_packed_kernel<1>);
    c.occ_packed[1] = occupancy(sb::scan_packed_kernel<2>);
    c.occ_packed[2] = occupancy(sb::scan_packed_kernel<4>);
    c.occ_packed[3] = occupancy(sb::scan_packed_kernel<8>);
    c.occ_packed_half[0] = occupancy(sb::scan_packed_half_kernel<1>);
    c.occ_packed_half[1] = occupancy(sb::scan_packed_half_kernel<2>);
    c.occ_pair[0] = occupancy(sb::scan_block_kernel<2, 3, kU>);
    c.occ_pair[1] = occupancy(sb::scan_block_kernel<2, 4, kU>);
    c.occ_pair[2] = occupancy(sb::scan_block_kernel<2, 5, kU>);
    c.occ_pair[3] = occupancy(sb::scan_block_kernel<2, 6, kU>);
    c.occ_pair[4] = occupancy(sb::scan_block_kernel<2, 7, kU>);
    c.occ_pair[5] = occupancy(sb::scan_block_kernel<2, 8, kU>);
    c.occ_block[block_index(4, 4)] = occupancy(sb::scan_block_kernel<4, 4, kU>);
    c.occ_block[block_index(8, 8)] = occupancy(sb::scan_block_kernel<8, 8, kU>);
    c.occ_block[block_index(16, 16)] = occupancy(sb::scan_block_kernel<16, 16, kU>);
    c.occ_block[block_index(32, 32)] = occupancy(sb::scan_block_kernel<32, 32, kU>);
    c.occ_block[block_index(4, 8)] = occupancy(sb::scan_block_kernel<4, 8, kU>);
    c.occ_block[block_index(4, 16)] = occupancy(sb::scan_block_kernel<4, 16, kU>);
    c.occ_block[block_index(8, 16)] = occupancy(sb::scan_block_kernel<8, 16, kU>);
#define SB_WIN_OCC(NW, OFF)                                                                       \
    c.occ_window[window_index(NW, OFF, false)] = occupancy(sb::scan_window_kernel<NW, OFF, false, kU>); \
    c.occ_window[window_index(NW, OFF, true)] = occupancy(sb::scan_window_kernel<NW, OFF, true, kU>);
    SB_WIN_OCC(2, 1) SB_WIN_OCC(2, 2) SB_WIN_OCC(2, 3) SB_WIN_OCC(2, 4)
    SB_WIN_OCC(3, 1) SB_WIN_OCC(3, 2) SB_WIN_OCC(3, 3) SB_WIN_OCC(3, 4)
    SB_WIN_OCC(4, 1) SB_WIN_OCC(4, 2) SB_WIN_OCC(4, 3) SB_WIN_OCC(4, 4)
#undef SB_WIN_OCC
#define SB_OCC_SWAR(M)                                                        \
    c.occ_swar[(M - 1) * 2] = occupancy(sb::scan_swar_kernel<M, false>); \
    c.occ_swar[(M - 1) * 2 + 1] = occupancy(sb::scan_swar_kernel<M, true>);
    SB_OCC_SWAR(1) SB_OCC_SWAR(2) SB_OCC_SWAR(3) SB_OCC_SWAR(4)
#undef SB_OCC_SWAR
    c.ready = true;
    return &c;
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Shard {
    int dev = 0;
    uint8_t *d_alloc = nullptr;   // cudaMalloc'ed block: kPadBytes of zeros, then the text (d_text)
    uint8_t *d_text = nullptr;
    size_t cap = 0;       // allocated bytes (group multiple + one spare group, zero padded)
    size_t base = 0;      // global offset of local byte 0
    size_t own = 0;       // start positions [base, base+own) belong to this shard
    size_t len = 0;       // valid bytes in the buffer (own + halo, clipped to the text end)
    uint32_t *d_packed = nullptr;  // 2-bit recoding, 16 characters per word (texts over <= 4 byte values), or nullptr
    size_t packed_bytes = 0;       // allocated bytes of d_packed (zero padded like the text)
};

/* Selectivity estimate per text byte from a byte histogram.  For a uniform alphabet both terms are log2(sigma).
 * The collision entropy H2 = -log2(sum p_c^2) gives the MEAN rate at which a text gram equals a pattern gram
 * (patterns are cut from the text); the min-entropy -log2(max p_c) bounds the worst pattern (all its bytes the
 * most frequent symbol) and also stands in for the inter-byte dependence of natural text that a byte
 * histogram cannot see.  The estimate is their mean: English text gives ~3.3 bits (H2 4.1, Hmin 2.5), so a
 * 4-byte fingerprint (13 bits) is judged too weak there and the 8-byte window is used instead. */
double collision_bits(const unsigned long long *hist)
{
    double tot = 0, s2 = 0, pmax = 0;
    for (int c = 0; c < 256; c++) tot += (double)hist[c];
    if (tot <= 0) return 8.0;
    for (int c = 0; c < 256; c++) {
        const double p = (double)hist[c] / tot;
        s2 += p * p;
        if (p > pmax) pmax = p;
    }
    double bits = 0.5 * (-log2(s2) - log2(pmax));
    if (bits < 0.05) bits = 0.05;
    return bits > 8.0 ? 8.0 : bits;
}

/* relative frequencies + a signature of them (texts with the same statistics share cached plans) */
void freq_from_hist(const double *hist, float *freq, uint64_t *sig)
{
    double tot = 0;
    for (int c = 0; c < 256; c++) tot += hist[c];
    uint64_t h = 0xCBF29CE484222325ull;
    for (int c = 0; c < 256; c++) {
        freq[c] = tot > 0 ? (float)(hist[c] / tot) : 1.0f / 256;
        uint32_t bits32;
        memcpy(&bits32, &freq[c], 4);
        h = (h ^ bits32) * 0x100000001B3ull;
    }
    *sig = h ? h : 1;
}

uint32_t load_le32(const unsigned char *p, int avail)
{
    uint32_t w = 0;
    for (int i = 0; i < 4 && i < avail; i++) w |= (uint32_t)p[i] << (8 * i);
    return w;
}

uint32_t fingerprint(const unsigned char *p, int B)
{
    uint32_t w[8];
    for (int i = 0; i < B / 4; i++) w[i] = load_le32(p + 4 * i, 4);
    switch (B) {
    case 4: return w[0];
    case 8: return sb::fp8(w[0], w[1]);
    case 16: return sb::fp16(w[0], w[1], w[2], w[3]);
    default: return sb::fp32(sb::fp16(w[0], w[1], w[2], w[3]), sb::fp16(w[4], w[5], w[6], w[7]));
    }
}

uint32_t fingerprint_pair(const unsigned char *p, int L)
{
    const uint32_t lo = load_le32(p, L < 4 ? L : 4), hi = L > 4 ? load_le32(p + 4, L - 4) : 0u;
    switch (L) {
    case 3: return sb::pair_fp<3>(lo, hi);
    case 4: return sb::pair_fp<4>(lo, hi);
    case 5: return sb::pair_fp<5>(lo, hi);
    case 6: return sb::pair_fp<6>(lo, hi);
    case 7: return sb::pair_fp<7>(lo, hi);
    default: return sb::pair_fp<8>(lo, hi);
    }
}

/* Which kernel scans a pattern of length m over a text with `bits` of selectivity per byte (see
 * collision_bits).  Returns the filter kind and, for scan_block, the fingerprint length L in *flen:
 *   0 = scan_swar (exact, m <= 4), 1/2/3 = scan_window (8/12/16-byte window at every position),
 *   4/8/16/32 = scan_block<B,L>: B-aligned windows of L >= B bytes, needs m >= L+B-1;
 *   20 = scan_block<2,L>: windows of L = 4..8 bytes at every even position, needs m >= L+1.
 * The block kernels are the fast ones (one compare per start position, no shifts) but an L-byte fingerprint
 * lets ~2^-(L*bits) of all start positions through to verification: pick the shortest L that carries
 * >= 20 bits (with the widest alignment B <= L that still fits the pattern), else the longest feasible L if it
 * carries >= 14 bits (or is 32), else a per-position window. */
int choose_filter(int m, unsigned flags, double bits, int *flen, bool periodic = false)
{
    static const int variants[7][2] = {{4, 4}, {8, 8}, {4, 8}, {16, 16}, {8, 16}, {4, 16}, {32, 32}};  // by L, widest B first
    *flen = 0;
    if ((flags & SB_FLAG_FILTER_B2) && m >= 4) return *flen = (m - 1 < 8 ? m - 1 : 8), kFilterPair;
    if ((flags & SB_FLAG_FILTER_B32) && m >= 63) return *flen = 32;
    if ((flags & SB_FLAG_FILTER_B8L16) && m >= 23) return *flen = 16, 8;
    if ((flags & SB_FLAG_FILTER_B4L16) && m >= 19) return *flen = 16, 4;
    if ((flags & SB_FLAG_FILTER_B4L8) && m >= 11) return *flen = 8, 4;
    if ((flags & SB_FLAG_FILTER_B16) && m >= 31) return *flen = 16;
    if ((flags & SB_FLAG_FILTER_B8) && m >= 15) return *flen = 8;
    if ((flags & SB_FLAG_FILTER_B4) && m >= 7) return *flen = 4;
    if (m == 4 && 3 * bits >= 15.0 && !(flags & (SB_FLAG_DENSE | SB_FLAG_FILTER_SHIFT))) return *flen = 3, kFilterPair;  // rare 4-byte pattern
    if (m <= 4) return kFilterShort;
    if (flags & SB_FLAG_FILTER_SHIFT) return kFilterWindow8;  // 8-byte window (exact for m <= 8)
    int bestB = 0, bestL = 0;
    for (int v = 0; v < 7; v++) {
        const int B = variants[v][0], L = variants[v][1];
        if (L + B - 1 > m || L <= bestL) continue;  // infeasible, or a wider alignment of this L was taken
        bestB = B;
        bestL = L;
        if (L * bits >= 20.0) return *flen = L, B;
    }
    if (bestL && (bestL * bits >= 14.0 || bestL > 16)) return *flen = bestL, bestB;
    /* a periodic pattern (period <= m/2) too long for an exact window: the block kernels count its dense runs
     * arithmetically (extend_run) and pre-check 32 candidates at a time, the per-position window kernels would
     * verify every position of a run lane by lane */
    if (bestL && periodic && m >= 17) return *flen = bestL, bestB;
    /* nearly constant texts (a 16-byte window carries < 8 bits): every position passes any fingerprint; the block
     * kernels probe 32 consecutive candidates per load instruction, the window kernels verify lane by lane */
    if (bestL && m >= 17 && 16.0 * bits < 8.0) return *flen = bestL, bestB;
    // 5 <= m < 11 (or 4 bytes too weak): windows of min(8, m-1) bytes at every even position
    if (m >= 5) {
        const int L = m - 1 < 8 ? m - 1 : 8;
        if (L * bits >= 15.0) return *flen = L, kFilterPair;  // measured: at 14 bits (sigma=4, m=8) the window kernel still wins
    }
    // per-position window of 2, 3 or 4 words: the narrowest that carries >= 14 bits, else the widest
    // that fits (a window never exceeds the pattern: min(m, 4*NW) bytes)
    for (int nw = 2; nw <= 4; nw++) {
        const int bytes = m < 4 * nw ? m : 4 * nw;
        if (bytes * bits >= 14.0 || bytes == m || nw == 4) return kFilterWindow8 + (nw - 2);
    }
    return kFilterWindow8;
}

/* Smallest sampling stride for which skipping sectors beats streaming: measured on B200, sampling one
 * sector in 2 or 4 (stride 64/128) is no faster than reading everything (DRAM delivers 64 bytes per
 * sampled sector even with the L2::64B hint, 128 without), from stride 160 on it wins (8.5 vs 7.1 TB/s).
 * SMART_B200_SKIP_MIN_STRIDE overrides (tuning). */
int skip_min_stride()
{
    static int v = 0;
    if (!v) {
        const char *e = getenv("SMART_B200_SKIP_MIN_STRIDE");
        v = e ? atoi(e) : 160;
        if (v < 32) v = 32;
    }
    return v;
}

/* below this many bits (m * bits) a pattern that fits in the window is counted in registers (dense mode)
 * instead of OR-reducing the compares; SMART_B200_DENSE_BITS overrides (tuning) */
double dense_bits()
{
    static double v = -1;
    if (v < 0) {
        const char *e = getenv("SMART_B200_DENSE_BITS");
        v = e ? atof(e) : 12.5;  // measured crossover (sigma 2 and 4, m = 5..16): dense wins below ~12.5 bits
    }
    return v;
}

/* SMART_B200_TMA=1 routes the disjoint-block filters through the TMA/shared-memory variant (experiment) */
bool use_tma()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SMART_B200_TMA");
        v = (e && atoi(e) > 0) ? 1 : 0;
    }
    return v == 1;
}

/* what make_plan needs to know about a resident text's 2-bit recoding */
struct PackInfo {
    bool packed = false;
    int bits = 0;   // bits per character of the recoding: 1 (two symbols), 2 (up to four) or 4 (up to sixteen)
    int nsym = 0;
    unsigned char sym[16] = {0};  // byte value of code c
};

struct Plan {
    int filter;
    int flen;           // scan_block: fingerprint length L (>= filter = B)
    int skip_grams;     // scan_skip: entries of the fingerprint list (= stride), 0 otherwise
    uint32_t list[4224]; // scan_skip: fingerprints of the pattern grams
    bool sparse;        // scan_swar: a match of the whole pattern is rare in this text
    int packed_half;    // scan_packed_half: sampling period in half-words (1, 2), 0 = word variants
    int packed_K;       // scan_packed: sampling period in words; 0 with filter == kFilterPacked: pattern outside the alphabet
    int pswar_bits;     // scan_pswar: bits per character of the recoding (1, 2); flen = KF, the chain length
    bool pswar_exact;   // scan_pswar: KF == m, the chain decides, nothing to verify
    sb::ScanArgs args;  // text/base/hi/ngroups/counters/pos filled per shard
};

void make_plan(Plan &pl, const sb_algo_desc *desc, const unsigned char *x, int m, unsigned flags, double bits,
               const PackInfo *pk = nullptr, const float *freq = nullptr)
{
    memset(&pl.args, 0, sizeof pl.args);
    static thread_local int border[SB_XSIZE + 1];
    sb_prep_mp(x, m, border);  // the border table (mp.c:26-35): the pattern's smallest period is m - border[m]
    const int period = m - border[m];
    pl.filter = choose_filter(m, flags, bits, &pl.flen, period * 2 <= m && !(flags & SB_FLAG_NO_PERIOD));
    pl.sparse = m * bits >= dense_bits() && !(flags & SB_FLAG_DENSE);  // scan_swar; scan_window decides below
    pl.skip_grams = 0;
    pl.packed_K = 0;
    pl.packed_half = 0;
    pl.pswar_bits = 0;
    pl.pswar_exact = false;
    sb::ScanArgs &a = pl.args;
    a.m = m;
    a.nwords = (m + 3) / 4;
    a.tailmask = (m & 3) ? ((1u << (8 * (m & 3))) - 1u) : 0xffffffffu;
    /* used on the device by the period-aware verification and the run counting of scan_block when the pattern is
     * at least two periods long */
    a.period = (period * 2 <= m && !(flags & SB_FLAG_NO_PERIOD)) ? period : 0;
    {   /* probe word: the pattern word with the byte that is rarest in the text (ties: the first one) */
        int r = m - 1;
        if (freq) {
            float best = 2.0f;
            for (int i = 0; i < m; i++)
                if (freq[x[i]] < best) {
                    best = freq[x[i]];
                    r = i;
                }
        }
        a.probe = r / 4;
        /* no byte of the pattern is rare in this text (>= 1/32 of the samples): the third probe word would reject
         * next to nothing, the pre-check keeps to the first and the last word */
        if (freq && freq[x[r]] >= 1.0f / 32) a.probe = -1;
    }
    /* Sub-linear sampling for the skip families: one G-byte gram every `stride` bytes (a multiple of the
     * 32-byte DRAM sector), G wide enough that a text gram rarely equals one of `stride` pattern grams. */
    const unsigned forced = SB_FLAG_FILTER_SHIFT | SB_FLAG_FILTER_B4 | SB_FLAG_FILTER_B8 | SB_FLAG_FILTER_B16 |
                            SB_FLAG_FILTER_B32 | SB_FLAG_FILTER_B4L8 | SB_FLAG_FILTER_B4L16 | SB_FLAG_FILTER_B8L16 |
                            SB_FLAG_FILTER_B2 | SB_FLAG_NO_SKIP;
    const bool can_pack = pk && pk->packed && !(flags & (SB_FLAG_NO_PACK | SB_FLAG_DENSE));
    if (desc->anchor == SB_ANCHOR_SUFFIX && !(flags & forced) && !(can_pack && m < (pk->bits == 1 ? kPackedBeatsSkipBelow1 : pk->bits == 2 ? kPackedBeatsSkipBelow2 : kPackedBeatsSkipBelow4))) {
        for (int G = 8; G <= 32; G *= 2) {
            const int stride = ((m - (G - 1)) / 32) * 32;
            if (stride < skip_min_stride()) break;
            if (G * bits < 14.0 + log2((double)stride)) continue;  // P(sampled gram is a pattern gram) <= ~1e-4
            const int anchor = m - (stride + G - 1);
            for (int d = 0; d < stride; d++) pl.list[d] = fingerprint(x + anchor + d, G);
            /* A (nearly) periodic pattern has few DISTINCT grams: every sampled hit would then fan out into
             * hundreds of candidate alignments resolved by one warp (measured: (ab)^n, m=4096 -> 0.03 GB/s).
             * Such patterns stay on the read-every-byte kernels, whose candidates spread over all warps. */
            {
                uint32_t sorted[4224];
                memcpy(sorted, pl.list, (size_t)stride * sizeof(uint32_t));
                std::sort(sorted, sorted + stride);
                int run = 1, worst = 1;
                for (int d = 1; d < stride; d++) {
                    run = sorted[d] == sorted[d - 1] ? run + 1 : 1;
                    worst = std::max(worst, run);
                }
                if (worst > 8) break;
            }
            pl.filter = kFilterSkip + G;
            pl.skip_grams = stride;
            a.stride = stride;
            a.anchor = anchor;
            return;
        }
    }
    /* Texts over <= 4 byte values: scan the 2-bit recoding (a quarter of the bytes).  Every K-th fingerprint (a
     * packed word or half-word) is looked up among those packed from x[anchor+d ..], d < span.
     * CPW = 16 characters per word (32 for two-symbol texts, 1 bit each).  m >= 2*CPW-1: whole words, every K-th
     * one, K = 1, 2, 4, 8 with m >= CPW*(K+1)-1; CPW-1 <= m < 2*CPW-1: half-words (only if CPW/2 characters carry
     * >= 14 bits), both halves or the low one. */
    /* Short patterns on a 1- or 2-bit recoding: the SWAR chain of scan_pswar, 32 / 16 start positions per
     * instruction.  The chain holds KF pattern characters: all m of them when that fits one word and fewer than 20
     * bits of selectivity are at stake otherwise (EXACT: nothing to verify, dense matches are counted in
     * registers), else the ~20-bit prefix/suffix of the pattern as a filter in front of the byte-wise verification. */
    if (pk && pk->packed && pk->bits <= 2 && m <= (pk->bits == 1 ? kPswarMaxM1 : kPswarMaxM2) &&
        !(flags & (forced | SB_FLAG_NO_PACK | SB_FLAG_DENSE))) {
        unsigned char code[256];
        memset(code, 0xff, sizeof code);
        for (int c = pk->nsym - 1; c >= 0; c--) code[pk->sym[c]] = (unsigned char)c;
        bool inside = true;
        for (int i = 0; i < m && inside; i++) inside = code[x[i]] != 0xff;
        if (!inside) {  // a byte the text does not contain: no occurrence, nothing to scan
            pl.filter = kFilterPacked;
            pl.packed_K = 0;
            return;
        }
        const int chain_max = 32 / pk->bits;
        int kf = (int)ceil(20.0 / (bits > 0.05 ? bits : 0.05) - 0.05);  // (a sampled histogram gives 0.999 for 1 bit)
        if (kf > chain_max) kf = chain_max;
        if (kf >= m || m <= 4) kf = m <= chain_max ? m : chain_max;
        pl.filter = kFilterPswar;
        pl.flen = kf;
        pl.pswar_bits = pk->bits;
        pl.pswar_exact = kf == m;
        a.cpw = chain_max;
        a.stride = kf;
        a.anchor = (desc->anchor == SB_ANCHOR_SUFFIX) ? m - kf : 0;
        const uint32_t unit = pk->bits == 1 ? 0xffffffffu : 0x55555555u;
        for (int k = 0; k < kf; k++) a.H[k] = (uint32_t)code[x[a.anchor + k]] * unit;
        return;
    }
    const int cpw = pk && pk->packed ? 32 / pk->bits : 16;  // characters per packed word
    /* a 4-bit copy only halves the bytes: it pays from K = 2 on (m >= 23), not for K = 1 or half-words, whose
     * probe-bound rates (3.6 / 2-3.4 TB/s of packed bytes) would then be no better than the byte kernels */
    const bool wide = pk && pk->packed && pk->bits == 4;
    const bool half_ok = !wide && m >= cpw - 1 && m < 2 * cpw - 1 && (cpw / 2) * bits >= 14.0;
    const bool word_ok = m >= (wide ? 3 * cpw - 1 : 2 * cpw - 1);
    if (pk && pk->packed && (word_ok || half_ok) && !(flags & (forced | SB_FLAG_NO_PACK | SB_FLAG_DENSE))) {
        const int gram = word_ok ? cpw : cpw / 2;           // characters per fingerprint
        int K = gram == cpw ? 8 : 2;                        // sampling period in fingerprints
        while (gram * K + gram - 1 > m) K >>= 1;
        const int span = gram * K;
        const int anchor = desc->anchor == SB_ANCHOR_SUFFIX ? m - (span + gram - 1) : 0;
        unsigned char code[256];
        memset(code, 0xff, sizeof code);
        for (int c = pk->nsym - 1; c >= 0; c--) code[pk->sym[c]] = (unsigned char)c;
        bool inside = true;
        for (int i = 0; i < m && inside; i++) inside = code[x[i]] != 0xff;
        if (!inside) {  // a byte the text does not contain: no occurrence, nothing to scan
            pl.filter = kFilterPacked;
            pl.packed_K = 0;
            return;
        }
        for (int d = 0; d < span; d++) {
            uint32_t w = 0;
            for (int k = 0; k < gram; k++) w |= (uint32_t)code[x[anchor + d + k]] << (pk->bits * k);
            pl.list[d] = w;
        }
        uint32_t sorted[288];
        memcpy(sorted, pl.list, (size_t)span * sizeof(uint32_t));
        std::sort(sorted, sorted + span);
        int run = 1, worst = 1;
        for (int d = 1; d < span; d++) {
            run = sorted[d] == sorted[d - 1] ? run + 1 : 1;
            worst = std::max(worst, run);
        }
        if (worst <= 8) {  // periodic patterns stay on the byte kernels (same reason as for scan_skip above)
            pl.filter = kFilterPacked + (gram == cpw ? K : 10 + K);
            pl.packed_K = K;
            pl.packed_half = gram == cpw ? 0 : K;
            a.cpw = cpw;
            pl.skip_grams = span;
            a.stride = span;
            a.anchor = anchor;
            return;
        }
    }
    if (pl.filter == kFilterShort) {
        a.M0 = m == 4 ? 0xffffffffu : ((1u << (8 * m)) - 1u);
        a.X0 = load_le32(x, m) & a.M0;
    } else if (pl.filter <= kFilterWindow16) {
        const int nw = pl.filter + 1;                         // window words: 2, 3 or 4
        const int off = m >= 4 * nw ? 4 : m - 4 * (nw - 1);   // last word overlaps when the window is the pattern
        a.X0 = load_le32(x, 4);
        for (int k = 1; k < nw; k++) a.Xw[k - 1] = load_le32(x + 4 * (k - 1) + (k == nw - 1 ? off : 4), 4);
        a.X1 = a.Xw[nw - 2];
        if (nw == 2) a.G = sb::fp8(a.X0, a.X1);
        else if (nw == 3) a.G = sb::fp8(a.X0, a.Xw[0]) * sb::K2 + a.Xw[1];
        else a.G = sb::fp8(a.X0, a.Xw[0]) * sb::K2 + sb::fp8(a.Xw[1], a.Xw[2]);
        a.anchor = off;  // reused to carry OFF to launch_scan
        // when the window is the whole pattern and matches are frequent, count them in registers
        pl.sparse = !(m <= 4 * nw && (m * bits < dense_bits() || (flags & SB_FLAG_DENSE)));
    } else {
        const bool pair = pl.filter == kFilterPair;
        const int B = pair ? 2 : pl.filter, L = pl.flen;
        a.anchor = desc->anchor == SB_ANCHOR_SUFFIX ? m - (L + B - 1) : 0;
        for (int d = 0; d < B; d++) a.H[d] = pair ? fingerprint_pair(x + a.anchor + d, L) : fingerprint(x + a.anchor + d, L);
    }
}

/* what sb_last_timing reports: see include/smart_b200.h */
int filter_kind_of(const Plan &pl)
{
    if (pl.filter == kFilterPswar) return kFilterPswar + (pl.pswar_bits == 2 ? 40 : 0) + pl.flen;
    if (pl.filter == kFilterPair) return 20 + pl.flen;
    if (pl.filter >= kFilterPacked) return pl.filter;
    if (pl.filter >= 4 && pl.filter < kFilterSkip && pl.flen != pl.filter) return pl.filter * 100 + pl.flen;
    return pl.filter;
}

/* scan_pswar<BITS, KF, EXACT> instances: [KF][EXACT] */
using PswarFn = void (*)(const sb::ScanArgs);
#define SB_PSWAR_ROW(B, K) {sb::scan_pswar_kernel<B, K, false>, sb::scan_pswar_kernel<B, K, true>}
const PswarFn kPswar1[33][2] = {
    {nullptr, nullptr},   SB_PSWAR_ROW(1, 1),  SB_PSWAR_ROW(1, 2),  SB_PSWAR_ROW(1, 3),  SB_PSWAR_ROW(1, 4),  SB_PSWAR_ROW(1, 5),
    SB_PSWAR_ROW(1, 6),   SB_PSWAR_ROW(1, 7),  SB_PSWAR_ROW(1, 8),  SB_PSWAR_ROW(1, 9),  SB_PSWAR_ROW(1, 10), SB_PSWAR_ROW(1, 11),
    SB_PSWAR_ROW(1, 12),  SB_PSWAR_ROW(1, 13), SB_PSWAR_ROW(1, 14), SB_PSWAR_ROW(1, 15), SB_PSWAR_ROW(1, 16), SB_PSWAR_ROW(1, 17),
    SB_PSWAR_ROW(1, 18),  SB_PSWAR_ROW(1, 19), SB_PSWAR_ROW(1, 20), SB_PSWAR_ROW(1, 21), SB_PSWAR_ROW(1, 22), SB_PSWAR_ROW(1, 23),
    SB_PSWAR_ROW(1, 24),  SB_PSWAR_ROW(1, 25), SB_PSWAR_ROW(1, 26), SB_PSWAR_ROW(1, 27), SB_PSWAR_ROW(1, 28), SB_PSWAR_ROW(1, 29),
    SB_PSWAR_ROW(1, 30),  SB_PSWAR_ROW(1, 31), SB_PSWAR_ROW(1, 32)};
const PswarFn kPswar2[17][2] = {
    {nullptr, nullptr},   SB_PSWAR_ROW(2, 1),  SB_PSWAR_ROW(2, 2),  SB_PSWAR_ROW(2, 3),  SB_PSWAR_ROW(2, 4),  SB_PSWAR_ROW(2, 5),
    SB_PSWAR_ROW(2, 6),   SB_PSWAR_ROW(2, 7),  SB_PSWAR_ROW(2, 8),  SB_PSWAR_ROW(2, 9),  SB_PSWAR_ROW(2, 10), SB_PSWAR_ROW(2, 11),
    SB_PSWAR_ROW(2, 12),  SB_PSWAR_ROW(2, 13), SB_PSWAR_ROW(2, 14), SB_PSWAR_ROW(2, 15), SB_PSWAR_ROW(2, 16)};
#undef SB_PSWAR_ROW

/* launch the scan of one buffer on `stream`; the caller has set the device */
cudaError_t launch_scan(const DevCtx &c, const Plan &pl, sb::ScanArgs a, cudaStream_t stream)
{
    if (a.hi <= 0) return cudaSuccess;
    int per_sm;
    size_t group;
    const bool want_pos = a.pos != nullptr;
    if (pl.filter == kFilterPswar) {
        const int B = pl.pswar_bits, KF = pl.flen, ex = pl.pswar_exact ? 1 : 0;
        const PswarFn fn = B == 1 ? kPswar1[KF][ex] : kPswar2[KF][ex];
        int &occ = c.occ_pswar[B - 1][KF][ex];
        if (!occ) {
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, sb::kThreads, 0) != cudaSuccess || nb < 1) nb = 1;
            occ = nb;
        }
        const unsigned long long end = (unsigned long long)a.hi + a.anchor + KF;  // characters the chain can touch
        const unsigned long long words = (end + a.cpw - 1) / a.cpw;
        static_assert(kGroupPacked == 4 * 1024, "scan_pswar / scan_packed / scan_swar step U = 4 KiB per warp");
        a.ngroups = (words * 4 + kGroupPacked - 1) / kGroupPacked;
        const unsigned long long want_ctas = (a.ngroups + (sb::kThreads / 32) - 1) / (sb::kThreads / 32);
        const unsigned g = (unsigned)std::max(1ull, std::min(want_ctas, (unsigned long long)c.sms * occ));
        fn<<<g, sb::kThreads, 0, stream>>>(a);
        return cudaGetLastError();
    }
    if (pl.filter >= kFilterPacked) {
        if (!pl.packed_K) return cudaSuccess;  // pattern outside the text's alphabet: the count stays 0
        const unsigned long long end = (unsigned long long)a.hi + a.anchor + a.stride + 31;  // characters that matter (upper bound)
        a.nsamples = (end + a.cpw - 1) / a.cpw;                                               // packed words
        a.ngroups = (a.nsamples * 4 + kGroupPacked - 1) / kGroupPacked;
        const int ki = pl.packed_K == 1 ? 0 : pl.packed_K == 2 ? 1 : pl.packed_K == 4 ? 2 : 3;
        const unsigned long long want_ctas = (a.ngroups + (sb::kThreads / 32) - 1) / (sb::kThreads / 32);
        const int occ = pl.packed_half ? c.occ_packed_half[pl.packed_half - 1] : c.occ_packed[ki];
        const unsigned g = (unsigned)std::max(1ull, std::min(want_ctas, (unsigned long long)c.sms * occ));
        if (pl.packed_half == 1) {
            sb::scan_packed_half_kernel<1><<<g, sb::kThreads, 0, stream>>>(a);
            return cudaGetLastError();
        }
        if (pl.packed_half == 2) {
            sb::scan_packed_half_kernel<2><<<g, sb::kThreads, 0, stream>>>(a);
            return cudaGetLastError();
        }
        switch (pl.packed_K) {
        case 1: sb::scan_packed_kernel<1><<<g, sb::kThreads, 0, stream>>>(a); break;
        case 2: sb::scan_packed_kernel<2><<<g, sb::kThreads, 0, stream>>>(a); break;
        case 4: sb::scan_packed_kernel<4><<<g, sb::kThreads, 0, stream>>>(a); break;
        default: sb::scan_packed_kernel<8><<<g, sb::kThreads, 0, stream>>>(a); break;
        }
        (void)want_pos;
        return cudaGetLastError();
    }
    if (pl.filter >= kFilterSkip) {
        const int G = pl.filter - kFilterSkip;
        a.nsamples = ((unsigned long long)a.hi + a.anchor + a.stride - 2) / a.stride + 1;
        const unsigned long long per_cta = 32ull * 8 * (sb::kThreads / 32);
        const unsigned long long want_ctas = (a.nsamples + per_cta - 1) / per_cta;
        const int occ = c.occ_skip[G == 8 ? 0 : G == 16 ? 1 : 2];
        const unsigned g = (unsigned)std::max(1ull, std::min(want_ctas, (unsigned long long)c.sms * occ));
        if (G == 8) sb::scan_skip_kernel<8><<<g, sb::kThreads, 0, stream>>>(a);
        else if (G == 16) sb::scan_skip_kernel<16><<<g, sb::kThreads, 0, stream>>>(a);
        else sb::scan_skip_kernel<32><<<g, sb::kThreads, 0, stream>>>(a);
        return cudaGetLastError();
    }
    const bool swar = pl.filter == kFilterShort;  // m <= 4: the SIMD-within-a-register kernel
    const bool swar_dense = !pl.sparse && !want_pos;
    if (swar) {
        a.period = 1;  // the kernel's multiplier-by-one (see scan_swar_kernel)
        group = kGroupPacked;
        a.ngroups = ((unsigned long long)a.hi + group - 1) / group;
        per_sm = c.occ_swar[(a.m - 1) * 2 + (swar_dense ? 1 : 0)];
    } else if (pl.filter <= kFilterWindow16) {
        group = kGroupLong;
        a.ngroups = ((unsigned long long)a.hi + group - 1) / group;
        const bool dense = !pl.sparse && !want_pos;
        per_sm = c.occ_window[window_index(pl.filter + 1, a.anchor, dense)];
    } else {
        group = kGroupLong;
        const bool pair = pl.filter == kFilterPair;
        const unsigned long long end = (unsigned long long)a.hi + a.anchor + (pair ? 2 : pl.filter) + pl.flen;
        a.ngroups = (end + group - 1) / group;
        per_sm = pair ? c.occ_pair[pl.flen - 3] : c.occ_block[block_index(pl.filter, pl.flen)];
    }
    unsigned long long want = (a.ngroups + (sb::kThreads / 32) - 1) / (sb::kThreads / 32);
    unsigned long long maxb = (unsigned long long)c.sms * per_sm;
    unsigned grid = (unsigned)std::max(1ull, std::min(want, maxb));
    if (swar) {
#define SB_SWAR_LAUNCH(M)                                                                      \
    if (a.m == M) {                                                                            \
        if (swar_dense) sb::scan_swar_kernel<M, true><<<grid, sb::kThreads, 0, stream>>>(a);   \
        else sb::scan_swar_kernel<M, false><<<grid, sb::kThreads, 0, stream>>>(a);             \
    }
        SB_SWAR_LAUNCH(1) SB_SWAR_LAUNCH(2) SB_SWAR_LAUNCH(3) SB_SWAR_LAUNCH(4)
#undef SB_SWAR_LAUNCH
    } else if (pl.filter <= kFilterWindow16) {
        const int off = a.anchor;
        const bool dense = !pl.sparse && !want_pos;
        a.anchor = 0;
        const int nw = pl.filter + 1;
#define SB_WIN_LAUNCH(NW, OFF)                                                                   \
    if (nw == NW && off == OFF) {                                                                \
        if (dense) sb::scan_window_kernel<NW, OFF, true, kU><<<grid, sb::kThreads, 0, stream>>>(a); \
        else sb::scan_window_kernel<NW, OFF, false, kU><<<grid, sb::kThreads, 0, stream>>>(a);      \
    }
        SB_WIN_LAUNCH(2, 1) SB_WIN_LAUNCH(2, 2) SB_WIN_LAUNCH(2, 3) SB_WIN_LAUNCH(2, 4)
        SB_WIN_LAUNCH(3, 1) SB_WIN_LAUNCH(3, 2) SB_WIN_LAUNCH(3, 3) SB_WIN_LAUNCH(3, 4)
        SB_WIN_LAUNCH(4, 1) SB_WIN_LAUNCH(4, 2) SB_WIN_LAUNCH(4, 3) SB_WIN_LAUNCH(4, 4)
#undef SB_WIN_LAUNCH
    } else if (pl.filter == kFilterPair) {
        switch (pl.flen) {
        case 3: sb::scan_block_kernel<2, 3, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        case 4: sb::scan_block_kernel<2, 4, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        case 5: sb::scan_block_kernel<2, 5, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        case 6: sb::scan_block_kernel<2, 6, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        case 7: sb::scan_block_kernel<2, 7, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        default: sb::scan_block_kernel<2, 8, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        }
    } else if (pl.flen == pl.filter && use_tma()) {
        // experiment: the same filter fed by TMA bulk copies through shared memory
        static const int stages = getenv("SMART_B200_TMA_STAGES") ? atoi(getenv("SMART_B200_TMA_STAGES")) : sb::kTmaStages;
        static const int ctas = getenv("SMART_B200_TMA_CTAS") ? atoi(getenv("SMART_B200_TMA_CTAS")) : 6;
        a.stride = stages;
        const size_t smem = (size_t)stages * sb::kTmaTile + 2 * stages * sizeof(uint64_t);
        const unsigned long long end = (unsigned long long)a.hi + a.anchor + 2 * pl.filter;
        a.nsamples = round_up(end, 4096);  // within the zero padding of the shard buffer (kPadBytes rule)
        a.ngroups = (a.nsamples + sb::kTmaTile - 1) / sb::kTmaTile;
        const unsigned g = (unsigned)std::max(1ull, std::min<unsigned long long>(a.ngroups, (unsigned long long)ctas * c.sms));
#define SB_TMA_LAUNCH(B)                                                                                         \
    {                                                                                                            \
        cudaFuncSetAttribute(sb::scan_block_tma_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        sb::scan_block_tma_kernel<B><<<g, sb::kTmaThreads, smem, stream>>>(a);                                   \
    }
        if (pl.filter == 4) SB_TMA_LAUNCH(4)
        else if (pl.filter == 8) SB_TMA_LAUNCH(8)
        else if (pl.filter == 16) SB_TMA_LAUNCH(16)
        else SB_TMA_LAUNCH(32)
#undef SB_TMA_LAUNCH
    } else {
        switch (block_index(pl.filter, pl.flen)) {
        case 0: sb::scan_block_kernel<4, 4, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        case 1: sb::scan_block_kernel<8, 8, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        case 2: sb::scan_block_kernel<16, 16, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        case 3: sb::scan_block_kernel<32, 32, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        case 4: sb::scan_block_kernel<4, 8, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        case 5: sb::scan_block_kernel<4, 16, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        default: sb::scan_block_kernel<8, 16, kU><<<grid, sb::kThreads, 0, stream>>>(a); break;
        }
    }
    return cudaGetLastError();
}

/* stage the pattern words of slot `slot` for upload; returns the bytes to copy (zero padded past m) */
size_t stage_pattern(DevCtx &c, int slot, const unsigned char *x, int m)
{
    const size_t bytes = round_up((size_t)m + 8, 16);
    unsigned char *h = reinterpret_cast<unsigned char *>(c.pat_host(slot));
    memcpy(h, x, (size_t)m);
    memset(h + m, 0, bytes - (size_t)m);
    return bytes;
}

/* Bloom bitmap + exact table {fingerprint, d+1} of a skip / packed plan, built in pinned memory */
void build_aux(uint32_t *h, const uint32_t *list, int n, bool half)
{
    memset(h, 0, kAuxWords * sizeof(uint32_t));
    uint32_t *table = h + sb::kBloomWords;
    for (int d = 0; d < n; d++) {
        if (half) {  // 16-bit fingerprints index the bitmap directly (scan_packed_half)
            h[list[d] >> 5] |= 1u << (list[d] & 31);
        } else {
            const uint32_t s1 = sb::bloom_slot(list[d]), s2 = sb::bloom_slot2(list[d]);
            h[s1 >> 5] |= 1u << (s1 & 31);
            h[s2 >> 5] |= 1u << (s2 & 31);
        }
        uint32_t slot = sb::table_slot(list[d]);
        while (table[2 * slot + 1]) slot = (slot + 1) & (sb::kSkipTableSize - 1);  // linear probing
        table[2 * slot] = list[d];
        table[2 * slot + 1] = (uint32_t)d + 1;
    }
}

/* least recently used aux slot of a device, ready to be refilled: its pinned staging is no longer being read
 * by a copy, and `up` (the stream that will overwrite the device block) waits for the last kernel reading it */
AuxSlot *aux_evict(DevCtx &c, cudaStream_t up)
{
    AuxSlot *best = nullptr;
    for (int i = 0; i < kAuxSlots; i++) {
        AuxSlot &s = c.aux[i];
        if (!s.used) {
            best = &s;
            break;
        }
        if (!best || s.stamp < best->stamp) best = &s;
    }
    AuxSlot &s = *best;
    if (!s.used) {
        if (cudaMalloc(&s.d, kAuxWords * sizeof(uint32_t)) != cudaSuccess ||
            cudaHostAlloc(&s.h, kAuxWords * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.staged, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.last_use, cudaEventDisableTiming) != cudaSuccess) {
            set_err("allocation of an aux block failed%s%s");
            return nullptr;
        }
        s.used = true;
    }
    s.valid = false;
    if (cudaEventSynchronize(s.staged) != cudaSuccess || cudaStreamWaitEvent(up, s.last_use, 0) != cudaSuccess) {
        set_err("aux block synchronisation failed%s%s");
        return nullptr;
    }
    s.stamp = ++c.aux_clock;
    return &s;
}

/* device block with the aux data of plan `pl`: the cached one, or built and uploaded on `up` */
AuxSlot *aux_acquire(DevCtx &c, const Plan &pl, cudaStream_t up)
{
    const bool half = pl.packed_half != 0;
    for (int i = 0; i < kAuxSlots; i++) {
        AuxSlot &s = c.aux[i];
        if (s.valid && s.half == half && (int)s.list.size() == pl.skip_grams &&
            memcmp(s.list.data(), pl.list, (size_t)pl.skip_grams * sizeof(uint32_t)) == 0) {
            s.stamp = ++c.aux_clock;
            return &s;
        }
    }
    AuxSlot *s = aux_evict(c, up);
    if (!s) return nullptr;
    build_aux(s->h, pl.list, pl.skip_grams, half);
    if (cudaMemcpyAsync(s->d, s->h, kAuxWords * sizeof(uint32_t), cudaMemcpyHostToDevice, up) != cudaSuccess ||
        cudaEventRecord(s->staged, up) != cudaSuccess) {
        set_err("aux block upload failed%s%s");
        return nullptr;
    }
    s->list.assign(pl.list, pl.list + pl.skip_grams);
    s->half = half;
    s->valid = true;
    return s;
}

/* Plans are cached per (scan direction, pattern, flags, text statistics): the driver's loop searches one pattern
 * with every selected algorithm (smart.c:312-345), and all algorithms of one scan direction share a plan. */
struct PlanEntry {
    bool valid = false;
    int anchor = 0, m = 0;
    unsigned flags = 0;
    double bits = 0;
    uint64_t sig = 0;   // signature of the text's byte histogram (the probe word depends on it)
    PackInfo pack;
    unsigned long long stamp = 0;
    unsigned char x[SB_XSIZE];
    Plan plan;
};
constexpr int kPlanCache = 8;
PlanEntry *g_plans = nullptr;
unsigned long long g_plan_clock = 0;
std::mutex g_plan_mu;

bool same_pack(const PackInfo &a, const PackInfo &b)
{
    return a.packed == b.packed && (!a.packed || (a.bits == b.bits && a.nsym == b.nsym && memcmp(a.sym, b.sym, sizeof a.sym) == 0));
}

void plan_get(Plan &out, const sb_algo_desc *desc, const unsigned char *x, int m, unsigned flags, double bits,
              const PackInfo *pk, const float *freq = nullptr, uint64_t sig = 0)
{
    static const PackInfo none;
    const PackInfo &pack = pk ? *pk : none;
    flags &= ~SB_FLAG_POSITIONS;
    std::lock_guard<std::mutex> lk(g_plan_mu);
    if (!g_plans) g_plans = new PlanEntry[kPlanCache];
    PlanEntry *lru = nullptr;  // an unused entry, else the least recently used one
    for (int i = 0; i < kPlanCache; i++) {
        PlanEntry &e = g_plans[i];
        if (e.valid && e.anchor == desc->anchor && e.m == m && e.flags == flags && e.bits == bits && e.sig == sig && same_pack(e.pack, pack) &&
            memcmp(e.x, x, (size_t)m) == 0) {
            e.stamp = ++g_plan_clock;
            out = e.plan;
            return;
        }
        if (!lru || (lru->valid && (!e.valid || e.stamp < lru->stamp))) lru = &e;
    }
    make_plan(lru->plan, desc, x, m, flags, bits, pk, freq);
    lru->valid = true;
    lru->sig = sig;
    lru->anchor = desc->anchor;
    lru->m = m;
    lru->flags = flags;
    lru->bits = bits;
    lru->pack = pack;
    memcpy(lru->x, x, (size_t)m);
    lru->stamp = ++g_plan_clock;
    out = lru->plan;
}

/* positions are produced window by window: a bitmap of kPosWindow start positions (256 MiB) at a time */
constexpr unsigned long long kPosWindow = 1ull << 31;
constexpr unsigned kPosBlocks = (unsigned)(kPosWindow / 32 / sb::kCompactWords);  // compaction CTAs per full window

/* position output buffers of a device: room for `need` sorted positions, the window bitmap and its scan arrays */
bool ensure_pos_capacity(DevCtx &c, size_t need)
{
    if (need < 1) need = 1;
    if (!c.d_posbits) {
        if (cudaMalloc(&c.d_posbits, kPosWindow / 8) != cudaSuccess || cudaMalloc(&c.d_blocksum, kPosBlocks * sizeof(unsigned)) != cudaSuccess ||
            cudaMalloc(&c.d_blockoff, kPosBlocks * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc(&c.d_written, sizeof(unsigned long long)) != cudaSuccess) {
            set_err("cudaMalloc of the position bitmap failed%s%s");
            return false;
        }
    }
    if (c.pos_cap >= need) return true;
    cudaStreamSynchronize(c.stream);
    if (c.d_pos) cudaFree(c.d_pos);
    c.d_pos = nullptr;
    c.pos_cap = 0;
    if (cudaMalloc(&c.d_pos, need * sizeof(unsigned long long)) != cudaSuccess) {
        set_err("cudaMalloc of the position buffer failed%s%s");
        return false;
    }
    c.pos_cap = need;
    return true;
}

}  // namespace

struct sb_text {
    size_t n = 0;
    int nshards = 0;
    double bits = 8.0;  // collision entropy per byte (min over shards), from a sampled histogram
    float freq[256];    // relative byte frequencies of the samples (all shards): the probe word of a plan
    uint64_t sig = 0;   // signature of freq (plan cache key)
    bool have_freq = false;
    bool present[256] = {false};  // byte values seen in the histogram samples (a guess, verified by pack2_kernel)
    PackInfo pack;      // the 2-bit recoding, when every byte of every shard is one of <= 4 values
    Shard sh[kMaxShards];
};

extern "C" {

const char *sb_last_error(void) { return g_err; }
const char *sb_version(void) { return "smart_b200 0.1 (sm_100a)"; }

int sb_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_err("cudaGetDeviceCount: %s%s", cudaGetErrorString(e));
        return -1;
    }
    return n;
}

void sb_last_timing(sb_timing *out)
{
    if (out) *out = g_timing;
}

int sb_stream_timer_begin(int device)
{
    DevCtx *c = get_ctx(device);
    if (!c) return -1;
    CU_TRY(cudaEventRecord(c->t0, c->stream), -1);
    return 0;
}

double sb_stream_timer_end_ms(int device)
{
    DevCtx *c = get_ctx(device);
    if (!c) return -1.0;
    float ms = 0;
    CU_TRY(cudaEventRecord(c->t1, c->stream), -1.0);
    CU_TRY(cudaEventSynchronize(c->t1), -1.0);
    CU_TRY(cudaEventElapsedTime(&ms, c->t0, c->t1), -1.0);
    return (double)ms;
}

void *sb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        set_err("cudaHostAlloc failed%s%s");
        return nullptr;
    }
    return p;
}

void sb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

void sb_textgen_host(unsigned char *dst, size_t off, size_t len, int sigma, uint64_t seed)
{
    size_t i = 0;
    while (i < len) {
        const size_t g = off + i;
        const unsigned long long w = sb::gen_word(seed, g >> 3, sigma);
        for (size_t k = g & 7; k < 8 && i < len; k++, i++) dst[i] = (unsigned char)(w >> (8 * k));
    }
}

int sb_shard_range(size_t n_total, int nparts, int part, size_t *base, size_t *own, size_t *len)
{
    if (nparts < 1 || part < 0 || part >= nparts) return -1;
    const size_t slice = round_up((n_total + nparts - 1) / nparts, kShardAlign);
    const size_t b = std::min(n_total, (size_t)part * slice);
    const size_t o = std::min(slice, n_total - b);
    if (base) *base = b;
    if (own) *own = o;
    if (len) *len = std::min(n_total - b, o + kHalo);
    return 0;
}

/* allocate the shards [first, first+count) of an nparts-way partition of an n-byte text */
static sb_text *text_alloc(size_t n, const int *devices, int nparts, int first, int count)
{
    if (nparts <= 0) nparts = 1;
    if (count > kMaxShards || first < 0 || first + count > nparts) {
        set_err("bad shard selection%s%s");
        return nullptr;
    }
    sb_text *t = new sb_text();
    t->n = n;
    t->nshards = count;
    for (int g = 0; g < count; g++) {
        Shard &s = t->sh[g];
        s.dev = devices ? devices[g] : g;
        sb_shard_range(n, nparts, first + g, &s.base, &s.own, &s.len);
        s.cap = round_up(s.len, sb::kPadBytes) + sb::kPadBytes;
        DevCtx *c = get_ctx(s.dev);
        if (!c) {
            sb_text_free(t);
            return nullptr;
        }
        /* a zeroed front pad: the slow paths read probe words of candidates that start up to SB_XSIZE + 31 bytes
         * BEFORE local byte 0 (start positions owned by the previous shard) without a bounds check */
        if (cudaMalloc(&s.d_alloc, s.cap + 2 * sb::kPadBytes) != cudaSuccess) {  // + a back pad the same reads may touch
            set_err("cudaMalloc of a text shard failed%s%s");
            s.d_alloc = nullptr;
            sb_text_free(t);
            return nullptr;
        }
        s.d_text = s.d_alloc + sb::kPadBytes;
        if (cudaMemsetAsync(s.d_alloc, 0, sb::kPadBytes, c->stream) != cudaSuccess) {
            set_err("cudaMemset of a text shard failed%s%s");
            sb_text_free(t);
            return nullptr;
        }
    }
    return t;
}

/* sampled byte histogram of every shard -> t->bits (min over shards) */
/* Texts over at most sixteen byte values also get a recoding at 4, 2 (<= 4 values) or 1 (two values) bits per
 * character for scan_packed (a half / a quarter / an eighth of the bytes to stream).  The alphabet is guessed from the histogram samples and VERIFIED by the packing pass over every byte;
 * a text that turns out to hold another value simply has no packed copy.  SMART_B200_NO_PACK=1 disables it.
 * Costs n/2, n/4 or n/8 bytes of HBM and one extra pass at upload/generation time. */
static bool text_try_pack(sb_text *t)
{
    static const bool off = getenv("SMART_B200_NO_PACK") && atoi(getenv("SMART_B200_NO_PACK")) > 0;
    int nsym = 0;
    unsigned char sym[16] = {0};
    for (int b = 0; b < 256; b++)
        if (t->present[b]) {
            if (nsym < 16) sym[nsym] = (unsigned char)b;
            nsym++;
        }
    if (off || nsym < 1 || nsym > 16 || t->n < (1u << 16)) return true;
    const int bits = nsym <= 2 ? 1 : nsym <= 4 ? 2 : 4;
    sb::PackLut lut;
    for (int c = 0; c < 16; c++) lut.sym[c] = c < nsym ? sym[c] : 0xffffffffu;  // unused codes match no byte
    bool ok = true;
    for (int g = 0; g < t->nshards && ok; g++) {
        Shard &s = t->sh[g];
        DevCtx *c = get_ctx(s.dev);
        if (!c) return false;
        cudaSetDevice(s.dev);
        s.packed_bytes = round_up((s.cap * bits + 7) / 8, sb::kPadBytes) + sb::kPadBytes;
        if (cudaMalloc(&s.d_packed, s.packed_bytes) != cudaSuccess) {
            cudaGetLastError();  // not fatal: the byte kernels need no packed copy
            s.d_packed = nullptr;
            ok = false;
            break;
        }
        unsigned bad = 0;
        if (cudaMemsetAsync(c->d_hist, 0, sizeof(unsigned), c->stream) != cudaSuccess) return false;
        if (bits == 1)
            sb::pack_kernel<1><<<c->sms * 8, 256, 0, c->stream>>>(s.d_text, s.len, s.cap, s.packed_bytes / 4, s.d_packed,
                                                                 lut, reinterpret_cast<unsigned *>(c->d_hist));
        else if (bits == 2)
            sb::pack_kernel<2><<<c->sms * 8, 256, 0, c->stream>>>(s.d_text, s.len, s.cap, s.packed_bytes / 4, s.d_packed,
                                                                 lut, reinterpret_cast<unsigned *>(c->d_hist));
        else
            sb::pack_kernel<4><<<c->sms * 8, 256, 0, c->stream>>>(s.d_text, s.len, s.cap, s.packed_bytes / 4, s.d_packed,
                                                                 lut, reinterpret_cast<unsigned *>(c->d_hist));
        if (cudaMemcpyAsync(&bad, c->d_hist, sizeof bad, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return false;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) return false;
        ok = bad == 0;
    }
    if (!ok) {  // a fifth byte value somewhere (or no memory): drop every packed copy
        for (int g = 0; g < t->nshards; g++)
            if (t->sh[g].d_packed) {
                cudaSetDevice(t->sh[g].dev);
                cudaFree(t->sh[g].d_packed);
                t->sh[g].d_packed = nullptr;
                t->sh[g].packed_bytes = 0;
            }
        return true;
    }
    t->pack.packed = true;
    t->pack.bits = bits;
    t->pack.nsym = nsym;
    memcpy(t->pack.sym, sym, sizeof sym);
    return true;
}

static bool text_measure_bits(sb_text *t)
{
    double bits = 8.0;
    double total[256] = {0};
    for (int g = 0; g < t->nshards; g++) {
        Shard &s = t->sh[g];
        if (s.len < 4096) continue;
        DevCtx *c = get_ctx(s.dev);
        if (!c) return false;
        const unsigned long long blocks = s.len / 4096;
        const unsigned long long nsamples = std::min<unsigned long long>(blocks, 4096);
        const unsigned long long stride = (blocks / nsamples) * 4096;
        unsigned long long hist[256];
        if (cudaMemsetAsync(c->d_hist, 0, sizeof hist, c->stream) != cudaSuccess) return false;
        sb::histogram_kernel<<<(unsigned)std::min<unsigned long long>(nsamples, 2048), 256, 0, c->stream>>>(
            s.d_text, s.len, nsamples, stride, c->d_hist);
        if (cudaMemcpyAsync(hist, c->d_hist, sizeof hist, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return false;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) return false;
        bits = std::min(bits, collision_bits(hist));
        for (int b = 0; b < 256; b++) {
            t->present[b] = t->present[b] || hist[b] != 0;
            total[b] += (double)hist[b];
        }
    }
    t->bits = bits;
    freq_from_hist(total, t->freq, &t->sig);
    t->have_freq = true;
    return text_try_pack(t);
}

static sb_text *text_fill_upload(sb_text *t, const unsigned char *host)
{
    if (!t) return nullptr;
    for (int g = 0; g < t->nshards; g++) {
        Shard &s = t->sh[g];
        DevCtx *c = get_ctx(s.dev);
        bool ok = c != nullptr;
        ok = ok && cudaMemsetAsync(s.d_text + s.len, 0, s.cap - s.len, c->stream) == cudaSuccess;
        ok = ok && (s.len == 0 ||
                    cudaMemcpyAsync(s.d_text, host + s.base, s.len, cudaMemcpyHostToDevice, c->stream) == cudaSuccess);
        if (!ok) {
            set_err("text upload failed: %s%s", cudaGetErrorString(cudaGetLastError()));
            sb_text_free(t);
            return nullptr;
        }
    }
    for (int g = 0; g < t->nshards; g++) {
        DevCtx *c = get_ctx(t->sh[g].dev);
        if (!c || cudaStreamSynchronize(c->stream) != cudaSuccess) {
            set_err("text upload sync failed%s%s");
            sb_text_free(t);
            return nullptr;
        }
    }
    if (!text_measure_bits(t)) {
        set_err("text statistics kernel failed: %s%s", cudaGetErrorString(cudaGetLastError()));
        sb_text_free(t);
        return nullptr;
    }
    return t;
}

static sb_text *text_fill_generate(sb_text *t, int sigma, uint64_t seed)
{
    if (!t) return nullptr;
    for (int g = 0; g < t->nshards; g++) {
        Shard &s = t->sh[g];
        DevCtx *c = get_ctx(s.dev);
        if (!c) {
            sb_text_free(t);
            return nullptr;
        }
        const unsigned long long nvec = s.cap / 16;
        // bytes past the shard's valid length must read as zero: generate up to base+len only
        sb::textgen_kernel<<<c->sms * 8, 256, 0, c->stream>>>(reinterpret_cast<uint4 *>(s.d_text), nvec, s.base,
                                                              s.base + s.len, seed, sigma);
        if (cudaGetLastError() != cudaSuccess) {
            set_err("textgen launch failed%s%s");
            sb_text_free(t);
            return nullptr;
        }
    }
    for (int g = 0; g < t->nshards; g++) {
        DevCtx *c = get_ctx(t->sh[g].dev);
        cudaError_t e = c ? cudaStreamSynchronize(c->stream) : cudaErrorUnknown;
        if (e != cudaSuccess) {
            set_err("textgen failed: %s%s", cudaGetErrorString(e));
            sb_text_free(t);
            return nullptr;
        }
    }
    if (!text_measure_bits(t)) {
        set_err("text statistics kernel failed: %s%s", cudaGetErrorString(cudaGetLastError()));
        sb_text_free(t);
        return nullptr;
    }
    return t;
}

sb_text *sb_text_upload(const unsigned char *host, size_t n, const int *devices, int ndev)
{
    if (!host && n) {
        set_err("null host text%s%s");
        return nullptr;
    }
    if (ndev <= 0) ndev = 1;
    return text_fill_upload(text_alloc(n, devices, ndev, 0, ndev), host);
}

sb_text *sb_text_upload_part(const unsigned char *host, size_t n_total, int nparts, int part, int device)
{
    if (!host && n_total) {
        set_err("null host text%s%s");
        return nullptr;
    }
    return text_fill_upload(text_alloc(n_total, &device, nparts, part, 1), host);
}

sb_text *sb_text_generate(size_t n, int sigma, uint64_t seed, const int *devices, int ndev)
{
    if (sigma < 1 || sigma > 256) {
        set_err("sigma out of range%s%s");
        return nullptr;
    }
    if (ndev <= 0) ndev = 1;
    return text_fill_generate(text_alloc(n, devices, ndev, 0, ndev), sigma, seed);
}

/* ---- SMART's own rand<sigma> buffers (src/textgen.c:34-54), generated on the device ---- */
namespace {
constexpr int kRandDeg = 31;  // r[i] = r[i-3] + r[i-31]: characteristic polynomial x^31 - x^28 - 1 over Z_2^32
struct RandPoly {
    uint32_t c[kRandDeg];
};
RandPoly rand_mul(const RandPoly &p, const RandPoly &q)
{
    uint32_t prod[2 * kRandDeg - 1] = {0};
    for (int i = 0; i < kRandDeg; i++)
        for (int j = 0; j < kRandDeg; j++) prod[i + j] += p.c[i] * q.c[j];
    for (int d = 2 * kRandDeg - 2; d >= kRandDeg; d--) {  // x^d = x^(d-3) + x^(d-31)
        prod[d - 3] += prod[d];
        prod[d - 31] += prod[d];
    }
    RandPoly r;
    memcpy(r.c, prod, sizeof r.c);
    return r;
}
RandPoly rand_xpow(unsigned long long n)
{
    RandPoly result = {{1}}, base = {{0, 1}};
    while (n) {
        if (n & 1) result = rand_mul(result, base);
        base = rand_mul(base, base);
        n >>= 1;
    }
    return result;
}
/* the 31 values a_0..a_30 the recurrence starts from: glibc srandom(1) (r[0] = 1, r[i] = 16807 r[i-1] mod 2^31-1 by
 * Schrage's method, r[31..33] = r[0..2]); a_n = r[n+3], and rand()'s k-th output is a_(k+341) >> 1 */
void rand_base(uint32_t *a)
{
    int32_t r[34];
    r[0] = 1;
    for (int i = 1; i < 31; i++) {
        const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    for (int n = 0; n < kRandDeg; n++) a[n] = (uint32_t)r[n + 3];
}
}  // namespace

/* 31 consecutive values a_P .. a_(P+30) of the stream, by jumping (no stepping from the start) */
static void rand_state_at(unsigned long long P, const RandPoly *xP, uint32_t *out)
{
    uint32_t base[kRandDeg];
    rand_base(base);
    RandPoly q = xP ? *xP : rand_xpow(P);
    for (int j = 0; j < kRandDeg; j++) {  // a_(P+j) = sum_k coeff_k(x^(P+j)) a_k
        uint32_t v = 0;
        for (int k = 0; k < kRandDeg; k++) v += q.c[k] * base[k];
        out[j] = v;
        const uint32_t top = q.c[kRandDeg - 1];  // multiply by x: x^31 = x^28 + 1
        for (int k = kRandDeg - 1; k > 0; k--) q.c[k] = q.c[k - 1];
        q.c[0] = top;
        q.c[28] += top;
    }
}

static int rand_file_index(int sigma)
{
    static const int sigmas[8] = {2, 4, 8, 16, 32, 64, 128, 250};  // the files textgen.c writes, in stream order
    for (int f = 0; f < 8; f++)
        if (sigmas[f] == sigma) return f;
    return -1;
}

int sb_textgen_rand_host(unsigned char *dst, int sigma, size_t off, size_t len)
{
    const int file = rand_file_index(sigma);
    if (file < 0 || !dst || off > 5000000 || len > 5000000 - off) return -1;
    uint32_t a[kRandDeg];
    rand_state_at(341ull + 5000000ull * (unsigned long long)file + off, nullptr, a);
    int head = 0;
    for (size_t i = 0; i < len; i++) {  // the same ring the device threads run (rand_textgen_kernel)
        const uint32_t out = a[head];
        dst[i] = (unsigned char)((out >> 1) % (uint32_t)sigma);
        const int k28 = head + 28 >= 31 ? head + 28 - 31 : head + 28;
        a[head] = a[k28] + out;
        head = head + 1 == 31 ? 0 : head + 1;
    }
    return 0;
}

sb_text *sb_text_generate_rand(int sigma, size_t n, const int *devices, int ndev)
{
    const int file = rand_file_index(sigma);
    if (file < 0 || n < 1 || n > 5000000) {
        set_err("sb_text_generate_rand: sigma must be one of SMART's rand buffers (2..128, 250) and n <= 5000000%s%s");
        return nullptr;
    }
    if (ndev <= 0) ndev = 1;
    sb_text *t = text_alloc(n, devices, ndev, 0, ndev);
    if (!t) return nullptr;
    const RandPoly xB = rand_xpow(sb::kRandBlock);
    for (int g = 0; g < t->nshards; g++) {
        Shard &s = t->sh[g];
        DevCtx *c = get_ctx(s.dev);
        if (!c) {
            sb_text_free(t);
            return nullptr;
        }
        const unsigned nblocks = (unsigned)((s.len + sb::kRandBlock - 1) / sb::kRandBlock);
        std::vector<uint32_t> states((size_t)nblocks * kRandDeg);
        // stream index of the shard's first byte: 5 000 000 outputs per earlier file, then the offset inside this one
        RandPoly p = rand_xpow(341ull + 5000000ull * (unsigned long long)file + s.base);
        for (unsigned b = 0; b < nblocks; b++) {
            rand_state_at(0, &p, states.data() + (size_t)b * kRandDeg);
            p = rand_mul(p, xB);
        }
        uint32_t *d_states = nullptr;
        bool ok = nblocks == 0 || cudaMalloc(&d_states, states.size() * sizeof(uint32_t)) == cudaSuccess;
        ok = ok && cudaMemsetAsync(s.d_text, 0, s.cap, c->stream) == cudaSuccess;
        if (ok && nblocks) {
            ok = cudaMemcpyAsync(d_states, states.data(), states.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream) ==
                 cudaSuccess;
            if (ok) sb::rand_textgen_kernel<<<(nblocks + 127) / 128, 128, 0, c->stream>>>(s.d_text, s.len, d_states, nblocks, sigma);
            ok = ok && cudaGetLastError() == cudaSuccess;
        }
        ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
        if (d_states) cudaFree(d_states);
        if (!ok) {
            set_err("sb_text_generate_rand failed: %s%s", cudaGetErrorString(cudaGetLastError()));
            sb_text_free(t);
            return nullptr;
        }
    }
    if (!text_measure_bits(t)) {
        set_err("text statistics kernel failed: %s%s", cudaGetErrorString(cudaGetLastError()));
        sb_text_free(t);
        return nullptr;
    }
    return t;
}

sb_text *sb_text_generate_part(size_t n_total, int sigma, uint64_t seed, int nparts, int part, int device)
{
    if (sigma < 1 || sigma > 256) {
        set_err("sigma out of range%s%s");
        return nullptr;
    }
    return text_fill_generate(text_alloc(n_total, &device, nparts, part, 1), sigma, seed);
}

size_t sb_text_length(const sb_text *t) { return t ? t->n : 0; }
int sb_text_shards(const sb_text *t) { return t ? t->nshards : 0; }
double sb_text_entropy_bits(const sb_text *t) { return t ? t->bits : 0.0; }
int sb_text_packed_bits(const sb_text *t) { return t && t->pack.packed ? t->pack.bits : 0; }

int sb_text_download(const sb_text *t, size_t off, size_t len, unsigned char *host)
{
    if (!t || off > t->n || len > t->n - off || (!host && len)) {
        set_err("bad range in sb_text_download%s%s");
        return -1;
    }
    size_t done = 0;
    while (done < len) {
        const size_t g_off = off + done;
        int g = 0;
        while (g + 1 < t->nshards && g_off >= t->sh[g].base + t->sh[g].own) g++;
        const Shard &s = t->sh[g];
        if (g_off < s.base || g_off >= s.base + s.len) {
            set_err("sb_text_download: range not held by this handle%s%s");
            return -1;
        }
        const size_t local = g_off - s.base;
        const size_t take = std::min(len - done, s.len - local);
        CU_TRY(cudaSetDevice(s.dev), -1);
        CU_TRY(cudaMemcpy(host + done, s.d_text + local, take, cudaMemcpyDeviceToHost), -1);
        done += take;
    }
    return 0;
}

void sb_text_free(sb_text *t)
{
    if (!t) return;
    for (int g = 0; g < t->nshards; g++)
        if (t->sh[g].d_alloc) {
            cudaSetDevice(t->sh[g].dev);
            cudaFree(t->sh[g].d_alloc);
            if (t->sh[g].d_packed) cudaFree(t->sh[g].d_packed);
        }
    delete t;
}

/* device used by the host-buffer entry points (sb_search_host, sb_plugin_search): sb_set_device(), else the
 * environment variable SMART_B200_DEVICE, else 0.  One process per GPU sets it to its local rank. */
static int g_host_device = -1;
static int host_device()
{
    if (g_host_device < 0) {
        const char *e = getenv("SMART_B200_DEVICE");
        g_host_device = e ? atoi(e) : 0;
    }
    return g_host_device;
}

int sb_set_device(int device)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        set_err("sb_set_device: no such CUDA device%s%s");
        return SB_ERR_ARG;
    }
    g_host_device = device;
    return 0;
}

int sb_get_device(void) { return host_device(); }

/* the reference's `return -1` guards: m below the algorithm's minimum (hash3.c:31, raita.c:38, qf23.c:44, ...)
 * or, for the double-window searches, a text shorter than two windows - where the window is the pattern already cut
 * to 16 characters (dbww.c:33-35: `if (m > 16) m = 16; if ((n/m) < 2) return -1;`) */
static bool outside_domain(const sb_algo_desc *d, int m, unsigned long long n)
{
    const unsigned long long window = m > 16 ? 16 : m;
    return m < d->min_m || (d->min_ratio > 0 && n / window < (unsigned long long)d->min_ratio);
}

/* ---- asynchronous searches: tickets ------------------------------------------------------------------
 * A ticket owns slot `idx` (pattern buffer, counters, events) on every device of its text.  submit enqueues
 *   ustream: pattern (+ aux block if not cached) H2D, event ev_up
 *   stream : wait ev_up, ev0, one scan kernel per shard, ev1
 *   dstream: wait ev1, counters D2H, ev_done
 * and returns; collect waits for ev_done only.  The scan stream therefore carries nothing but kernels, back to
 * back, while uploads and read-backs of neighbouring searches overlap them. */
struct Ticket {
    int state = 0;          // 0 free, 1 in flight, 2 finished at submit time (result in `immediate`)
    long long immediate = 0;
    int ndev = 0;
    int devs[kMaxShards];
    double prep_ms = 0, upload_ms = 0;
    int launched = 0, filter_kind = 0;
};
static Ticket g_tickets[kTickets];
static int g_ticket_cursor = 0;
static std::mutex g_ticket_mu;

static int ticket_alloc()
{
    std::lock_guard<std::mutex> lk(g_ticket_mu);
    for (int i = 0; i < kTickets; i++) {
        const int idx = (g_ticket_cursor + i) % kTickets;
        if (g_tickets[idx].state == 0) {
            g_tickets[idx] = Ticket();
            g_tickets[idx].state = 1;
            g_ticket_cursor = (idx + 1) % kTickets;
            return idx;
        }
    }
    return -1;
}

static void ticket_release(int idx)
{
    std::lock_guard<std::mutex> lk(g_ticket_mu);
    g_tickets[idx].state = 0;
}

static bool ticket_events(DevCtx &c, int idx)
{
    if (c.ev_done[idx]) return true;
    return cudaEventCreateWithFlags(&c.ev_up[idx], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreate(&c.ev0[idx]) == cudaSuccess && cudaEventCreate(&c.ev1[idx]) == cudaSuccess &&
           cudaEventCreateWithFlags(&c.ev_done[idx], cudaEventDisableTiming) == cudaSuccess;
}

/* per-shard kernel arguments of a plan */
static sb::ScanArgs shard_args(const Plan &pl, const sb_text *t, int g, int m)
{
    const Shard &s = t->sh[g];
    sb::ScanArgs a = pl.args;
    a.text = s.d_text;
    a.packed = s.d_packed;
    a.base = s.base;
    long long hi = (long long)t->n - m + 1 - (long long)s.base;
    if (hi > (long long)s.own) hi = (long long)s.own;
    a.hi = hi > 0 ? hi : 0;
    return a;
}

/* enqueue a count-only search; returns the ticket index or an SB_ERR_* code */
static long long submit_core(const sb_algo_desc *desc, const sb_text *t, const unsigned char *x, int m, unsigned flags)
{
    const int idx = ticket_alloc();
    if (idx < 0) {
        set_err("too many outstanding searches: collect some tickets first%s%s");
        return SB_ERR_ARG;
    }
    Ticket &tk = g_tickets[idx];
    if (outside_domain(desc, m, t->n)) {
        tk.state = 2;
        tk.immediate = SB_NOT_APPLICABLE;
        return idx;
    }
    /* ---- preprocessing stage: the plan (fingerprints, period, aux list) comes from the cache or is built ---- */
    const double t0 = now_ms();
    Plan pl;
    plan_get(pl, desc, x, m, flags, t->bits, &t->pack, t->have_freq ? t->freq : nullptr, t->have_freq ? t->sig : 0);
    const double t1 = now_ms();
    tk.filter_kind = filter_kind_of(pl);
    tk.prep_ms = t1 - t0;
    if ((long long)t->n - m + 1 <= 0 || (pl.filter == kFilterPacked && !pl.packed_K)) {
        tk.state = 2;  // pattern longer than the text, or with a byte the text does not contain: count 0
        tk.immediate = 0;
        return idx;
    }
#define SUBMIT_TRY(expr)                                                            \
    do {                                                                            \
        cudaError_t e_ = (expr);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            set_err("CUDA error in %s: %s", #expr, cudaGetErrorString(e_));         \
            ticket_release(idx);                                                    \
            return SB_ERR_CUDA;                                                     \
        }                                                                           \
    } while (0)
    for (int g = 0; g < t->nshards; g++) {
        bool first = true;  // shards may share a device (tests): one slot, one upload, one read-back per DEVICE
        for (int k = 0; k < g; k++) first = first && (t->sh[k].dev != t->sh[g].dev);
        if (!first) continue;
        DevCtx *c = get_ctx(t->sh[g].dev);
        if (!c || !ticket_events(*c, idx)) {
            ticket_release(idx);
            return SB_ERR_CUDA;
        }
        const size_t pat_bytes = stage_pattern(*c, idx, x, m);
        SUBMIT_TRY(cudaMemcpyAsync(c->pat_dev(idx), c->pat_host(idx), pat_bytes, cudaMemcpyHostToDevice, c->ustream));
        AuxSlot *aux = nullptr;
        if (pl.skip_grams) {
            aux = aux_acquire(*c, pl, c->ustream);
            if (!aux) {
                ticket_release(idx);
                return SB_ERR_CUDA;
            }
        }
        SUBMIT_TRY(cudaEventRecord(c->ev_up[idx], c->ustream));
        SUBMIT_TRY(cudaStreamWaitEvent(c->stream, c->ev_up[idx], 0));
        SUBMIT_TRY(cudaEventRecord(c->ev0[idx], c->stream));
        for (int k = g; k < t->nshards; k++) {
            if (t->sh[k].dev != t->sh[g].dev) continue;
            sb::ScanArgs a = shard_args(pl, t, k, m);
            if (a.hi <= 0) continue;
            a.pat = c->pat_dev(idx);
            a.counters = c->cnt_dev(idx);
            a.aux = aux ? aux->d : nullptr;
            SUBMIT_TRY(launch_scan(*c, pl, a, c->stream));
            tk.launched++;
        }
        if (aux) SUBMIT_TRY(cudaEventRecord(aux->last_use, c->stream));
        SUBMIT_TRY(cudaEventRecord(c->ev1[idx], c->stream));
        SUBMIT_TRY(cudaStreamWaitEvent(c->dstream, c->ev1[idx], 0));
        SUBMIT_TRY(cudaMemcpyAsync(c->cnt_host(idx), c->cnt_dev(idx), 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                   c->dstream));
        SUBMIT_TRY(cudaEventRecord(c->ev_done[idx], c->dstream));
        tk.devs[tk.ndev++] = t->sh[g].dev;
    }
#undef SUBMIT_TRY
    tk.upload_ms = now_ms() - t1;
    return idx;
}

/* wait for a ticket; fills g_timing (kernel_ms = max over devices of the CUDA-event time of its kernels) */
static long long collect_core(int idx)
{
    Ticket &tk = g_tickets[idx];
    memset(&g_timing, 0, sizeof g_timing);
    g_timing.prep_host_ms = tk.prep_ms;
    g_timing.upload_ms = tk.upload_ms;
    g_timing.filter_kind = tk.filter_kind;
    g_timing.kernels_launched = tk.launched;
    if (tk.state == 2) {
        const long long r = tk.immediate;
        ticket_release(idx);
        return r;
    }
    unsigned long long total = 0, verified = 0;
    double kernel_ms = 0;
    long long err = 0;
    for (int i = 0; i < tk.ndev; i++) {
        DevCtx *c = get_ctx(tk.devs[i]);
        cudaError_t e = c ? cudaEventSynchronize(c->ev_done[idx]) : cudaErrorUnknown;
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, c->ev0[idx], c->ev1[idx]);
        if (e != cudaSuccess) {
            set_err("CUDA error while collecting a search: %s%s", cudaGetErrorString(e));
            err = SB_ERR_CUDA;
            continue;
        }
        kernel_ms = std::max(kernel_ms, (double)ms);
        const unsigned long long *h = c->cnt_host(idx);
        total += h[0] - c->prev[idx][0];
        verified += h[2] - c->prev[idx][2];
        for (int k = 0; k < 4; k++) c->prev[idx][k] = h[k];
    }
    g_timing.kernel_ms = kernel_ms;
    g_timing.verified = verified;
    ticket_release(idx);
    return err ? err : (long long)total;
}

long long sb_search64_submit(int algo, const sb_text *t, const unsigned char *x, int m, unsigned flags)
{
    const sb_algo_desc *desc = sb_algo_desc_get(algo);
    if (!desc || !t || !x || m < 1 || m > SB_XSIZE || (flags & SB_FLAG_POSITIONS)) {
        set_err("bad argument to sb_search64_submit (algo/text/pattern/m, or SB_FLAG_POSITIONS)%s%s");
        return SB_ERR_ARG;
    }
    return submit_core(desc, t, x, m, flags);
}

long long sb_search_collect(long long ticket, double *pre_ms, double *run_ms)
{
    if (pre_ms) *pre_ms = 0;
    if (run_ms) *run_ms = 0;
    if (ticket < 0 || ticket >= kTickets || g_tickets[ticket].state == 0) {
        set_err("sb_search_collect: no such ticket%s%s");
        return SB_ERR_ARG;
    }
    const long long r = collect_core((int)ticket);
    if (pre_ms) *pre_ms = g_timing.prep_host_ms + g_timing.upload_ms;
    if (run_ms) *run_ms = g_timing.kernel_ms > 0 ? g_timing.kernel_ms : 1e-6;
    return r;
}

static long long search_positions(const sb_algo_desc *desc, const sb_text *t, const unsigned char *x, int m, unsigned flags,
                                  unsigned long long *positions, size_t cap, double *pre_ms, double *run_ms);

/* SB_FLAG_TABLE_SCAN: the reference algorithm's own table, built by prep.c, uploaded and used by its own scan loop on
 * the device (table_kernels.cuh).  Returns -100 when the family has no table-driven kernel (caller takes the default). */
static long long search_table(const sb_algo_desc *desc, const sb_text *t, const unsigned char *x, int m, double *pre_ms,
                              double *run_ms)
{
    int kind = 0;
    switch (desc->family) {
    case SB_FAM_KMP: kind = 601; break;
    case SB_FAM_MP: kind = 602; break;
    case SB_FAM_HOR: kind = 603; break;
    case SB_FAM_SO: kind = 604; break;
    case SB_FAM_SA: kind = 605; break;
    case SB_FAM_HASHQ: kind = 606; break;
    case SB_FAM_BM: kind = 607; break;
    case SB_FAM_QS: kind = strcmp(desc->name, "qs") == 0 ? 608 : strcmp(desc->name, "ssabs") == 0 ? 623 : 0; break;
    case SB_FAM_FS: kind = 624; break;
    case SB_FAM_GRASPM: kind = 625; break;
    case SB_FAM_TBM: kind = 626; break;
    case SB_FAM_AC: kind = 627; break;
    case SB_FAM_BADCHAR:
        if (strcmp(desc->name, "zt") == 0) kind = 628;
        else if (strcmp(desc->name, "iom") == 0) kind = 637;
        else if (strcmp(desc->name, "qlqs") == 0) kind = 638;
        else if (strcmp(desc->name, "tsw") == 0) kind = 639;
        else if (strcmp(desc->name, "tvsbs-w2") == 0) kind = 640;
        else if (strcmp(desc->name, "wc") == 0) kind = 647;
        else if (strcmp(desc->name, "wom") == 0) kind = 650;
        else if (strcmp(desc->name, "pbmh") == 0) kind = 651;
        else if (strcmp(desc->name, "fs-w1") == 0) kind = 648;
        break;
    case SB_FAM_SKIPQ: kind = 629; break;
    case SB_FAM_WFR:  // wfr (wfr.c), wfrq2..8 (wfrq4.c) and wfr2..8 (wfr4.c: wfrq's scan over wfr's table); not twfr*, lwfr*
        kind = (strncmp(desc->name, "wfr", 3) == 0 || strncmp(desc->name, "twfr", 4) == 0) ? 630 : 0;  // (twfr*: + the skip loop)
        break;
    case SB_FAM_RAITA: kind = 609; break;
    case SB_FAM_TUNEDBM: kind = 610; break;
    case SB_FAM_KR: kind = 611; break;
    case SB_FAM_QF: kind = 612; break;
    case SB_FAM_AUT: kind = 613; break;
    case SB_FAM_BNDM: kind = 617; break;
    case SB_FAM_BR: kind = strcmp(desc->name, "br") == 0 ? 618 : strcmp(desc->name, "tvsbs") == 0 ? 619 : 0; break;
    case SB_FAM_SMITH: kind = 620; break;
    case SB_FAM_SKIP: kind = 621; break;
    case SB_FAM_FJS: kind = 622; break;
    case SB_FAM_BF: kind = strcmp(desc->name, "ts") == 0 ? 634 : 0; break;  // (bf itself IS the default mode)
    case SB_FAM_KMPSKIP: kind = 635; break;
    case SB_FAM_BITPAR: kind = strncmp(desc->name, "aoso", 4) == 0 ? 645 : strcmp(desc->name, "svm0") == 0 ? 646 : 0; break;
    case SB_FAM_BNDMX:
        if (strcmp(desc->name, "sbndm2") == 0) kind = m <= SB_WORD ? 616 : 0;  // (m > 32: the default mode)
        else if (strncmp(desc->name, "bndmq", 5) == 0) kind = 631;
        else if (strcmp(desc->name, "fndm") == 0) kind = 632;
        else if (strcmp(desc->name, "sbndm") == 0) kind = 633;
        else if (strncmp(desc->name, "bxs", 3) == 0) kind = 641;
        else if (strcmp(desc->name, "sbndm-bmh") == 0) kind = 642;
        else if (strcmp(desc->name, "lbndm") == 0) kind = 643;
        break;
    case SB_FAM_FACTOR:
        if (strcmp(desc->name, "bom") == 0) kind = 614;
        else if (strcmp(desc->name, "ebom") == 0 || strcmp(desc->name, "sebom") == 0) kind = 615;  // sebom.c:61-77 is ebom.c:62-78
        else if (strcmp(desc->name, "bom2") == 0) kind = 636;
        else if (strncmp(desc->name, "bsdm", 4) == 0) kind = 644;
        else if (strcmp(desc->name, "fbom") == 0 || strcmp(desc->name, "sfbom") == 0) kind = 649;
        break;
    default: break;
    }
    if (!kind) return -100;
    const double t0 = now_ms();
    static thread_local int *table = nullptr;
    static thread_local int scratch[SB_XSIZE + 1];
    if (!table) table = (int *)malloc(kTableInts * sizeof(int));
    if (!table) return SB_ERR_NOMEM;
    int ntable = 256, sh1 = 0, qq = desc->q;
    unsigned lim = 0;
    switch (kind) {  // the preprocessing stage proper: bit-identical to preKmp / preMp / Pre_Horspool / preSo / preSA / hash3.c:37-55
    case 601:
        sb_prep_mp(x, m, scratch);
        sb_prep_kmp(x, m, scratch, table);
        ntable = m + 1;
        break;
    case 602:
        sb_prep_mp(x, m, table);
        ntable = m + 1;
        break;
    case 603: sb_prep_hbc(x, m, table); break;
    case 604: lim = sb_prep_so(x, m, reinterpret_cast<unsigned int *>(table)); break;
    case 605: lim = sb_prep_sa(x, m, reinterpret_cast<unsigned int *>(table)); break;
    case 606: sb_prep_hashq(x, m, desc->q, table, &sh1); break;
    case 607:  // bmBc[256] then bmGs[0..m)
        sb_prep_hbc(x, m, table);
        sb_prep_bmgs(x, m, scratch, table + 256);
        ntable = 256 + m;
        break;
    case 608: sb_prep_qsbc(x, m, table); break;
    case 609:
    case 610: sb_prep_hbc(x, m, table); break;  // preBmBc of raita.c:26-32 / tunedbm.c:26-32 == Pre_Horspool
    case 611: {
        int d = 0, hx = 0;
        sb_prep_kr(x, m, &d, &hx);
        lim = (unsigned)d;
        sh1 = hx;
        ntable = 1;
        table[0] = 0;
        break;
    }
    case 613:  // (m+1) x 256 16-bit states
        sb_prep_automaton(x, m, reinterpret_cast<unsigned short *>(table));
        ntable = ((m + 1) * 256 + 1) / 2;
        break;
    case 636:
    case 614: {  // (m+1) x 256 16-bit transitions, then m+1 terminal flags
        short *tr = reinterpret_cast<short *>(table);
        sb_prep_bom(x, m, tr, reinterpret_cast<unsigned char *>(tr + (size_t)(m + 1) * 256));
        ntable = ((m + 1) * 256 * 2 + (m + 1) + 3) / 4;
        break;
    }
    case 615: {  // the oracle of bom, then FT[256][256] (ebom.c:54-59)
        short *tr = reinterpret_cast<short *>(table);
        sb_prep_bom(x, m, tr, reinterpret_cast<unsigned char *>(tr + (size_t)(m + 1) * 256));
        const size_t ft_off = (((size_t)(m + 1) * 256 * 2 + (m + 1) + 3) / 4) * 2;  // in shorts, 4-byte aligned
        short *FT = tr + ft_off;
        for (int c1 = 0; c1 < 256; c1++) {
            const int q = tr[(size_t)m * 256 + c1];
            for (int c2 = 0; c2 < 256; c2++) FT[c1 * 256 + c2] = q >= 0 ? tr[(size_t)q * 256 + c2] : (short)-1;
        }
        lim = (unsigned)ft_off;
        ntable = (int)(ft_off / 2 + 32768);
        break;
    }
    case 649: {  // the oracle of bom, then the look-ahead table of fbom.c:53-64
        short *tr = reinterpret_cast<short *>(table);
        sb_prep_bom(x, m, tr, reinterpret_cast<unsigned char *>(tr + (size_t)(m + 1) * 256));
        const size_t ft_off = (((size_t)(m + 1) * 256 * 2 + (m + 1) + 3) / 4) * 2;  // in shorts, 4-byte aligned
        short *FT = tr + ft_off;
        for (int c1 = 0; c1 < 256; c1++) {
            const int q = tr[(size_t)m * 256 + c1];
            for (int c2 = 0; c2 < 256; c2++) {
                const int p = q >= 0 ? tr[(size_t)q * 256 + c2] : -1;
                FT[c1 * 256 + c2] = (short)(p >= 0 ? p : 2 * m + 1);
            }
        }
        const short q = tr[(size_t)m * 256 + x[m - 1]];
        for (int c1 = 0; c1 < 256; c1++) FT[c1 * 256 + x[m - 1]] = q;
        for (int c2 = 0; c2 < 256; c2++)
            if (FT[x[0] * 256 + c2] > m) FT[x[0] * 256 + c2] -= 1;
        lim = (unsigned)ft_off;
        ntable = (int)(ft_off / 2 + 32768);
        break;
    }
    case 616: sb_prep_sbndm2(x, m, reinterpret_cast<unsigned int *>(table), &sh1); break;
    case 617: sb_prep_bndm(x, m, reinterpret_cast<unsigned int *>(table)); break;
    case 623: sb_prep_qsbc(x, m, table); break;
    case 646: {  // svm0.c:44-52 over the first min(m, 31) characters
        const int k = m < SB_WORD - 1 ? m : SB_WORD - 1;
        const unsigned all = ~0u >> (SB_WORD - k);
        for (int c = 0; c < 256; c++) table[c] = (int)all;
        for (int j = k - 1, b = 0; j >= 0; j--, b++) table[x[j]] = (int)((unsigned)table[x[j]] & ~(1u << b));
        break;
    }
    case 647: {  // wc.c:41-70 FindBadPos (the reference keeps `max` in an int: the comparison truncates) + PreBadPosBC
        int last[256], pos = 0, mx = 1;
        double adv = 1;
        for (int c = 0; c < 256; c++) last[c] = -1;
        for (int i = 1; i <= m; i++) {
            const int gbc = i - last[x[i - 1]] - 1;
            adv = adv - (double)gbc * 0.5 + 1;
            last[x[i - 1]] = i - 1;
            if (adv > mx) {
                mx = (int)adv;
                pos = i;
            }
        }
        for (int c = 0; c < 256; c++) table[c] = pos + 1;
        for (int i = 0; i < pos; i++) table[x[i]] = pos - i;
        qq = pos;
        break;
    }
    case 650:
    case 651: {  // both rank characters by their count among the text's first 100 characters (wom.c:29-33, pbmh.c:33-34)
        unsigned char head[100] = {0};
        const size_t have = t->n < 100 ? t->n : 100;  // (the reference reads 100 bytes whatever n is)
        if (t->nshards > 0 && have > 0) {
            CU_TRY(cudaSetDevice(t->sh[0].dev), SB_ERR_CUDA);
            CU_TRY(cudaMemcpy(head, t->sh[0].d_text, have < t->sh[0].len ? have : t->sh[0].len, cudaMemcpyDeviceToHost),
                   SB_ERR_CUDA);
        }
        int cntc[256] = {0};
        for (int i = 0; i < 100; i++) cntc[head[i]]++;
        if (kind == 650) {  // wom.c:35-56 FindWorstOccurrence + PrecomputeWOH
            double freq[256], adv = 1.0, mx = 1.0;
            int lp[256], q = 0;
            for (int c = 0; c < 256; c++) {
                freq[c] = cntc[c] / 100.0;
                lp[c] = -1;
            }
            for (int i = 1; i <= m; i++) {
                const int gbc = i - lp[x[i - 1]] - 1;
                adv = adv - freq[x[i - 1]] * gbc + 1.0;
                lp[x[i - 1]] = i - 1;
                if (adv > mx) {
                    mx = adv;
                    q = i;
                }
            }
            for (int c = 0; c < 256; c++) table[c] = q + 1;
            for (int i = 0; i < q; i++) table[x[i]] = q - i;
            qq = q;
        } else {            // pbmh.c:36-47: positions by ascending frequency (the reference's bubble sort is stable)
            static thread_local int order[SB_XSIZE];
            for (int i = 0; i < m; i++) order[i] = i;
            std::stable_sort(order, order + m, [&](int u, int v) { return cntc[x[u]] < cntc[x[v]]; });
            sb_prep_hbc(x, m, table);
            for (int i = 0; i < m; i++) table[256 + i] = order[i];
            ntable = 256 + m;
        }
        break;
    }
    case 648:
    case 624:
        sb_prep_lastocc(x, m, table);
        sb_prep_fsgs(x, m, table + 256);
        ntable = 256 + m + 1;
        break;
    case 625:
        sb_prep_lastocc(x, m, table);
        sb_prep_graspm(x, m, table + 256, table + 512);
        ntable = 512 + m + 1;
        break;
    case 626:
        sb_prep_hbc(x, m, table);
        sb_prep_bmgs(x, m, scratch, table + 256);
        ntable = 256 + m;
        break;
    case 637: {  // iom.c:36-41: the look-ahead distance after a last-character match / mismatch, and their tables
        int step1 = 2 * m - 1, step2 = 2 * m - 1;
        for (int i = 0; i < m - 1; i++) {
            if (x[i] == x[m - 1]) step1 = 2 * m - i - 2;
            else step2 = 2 * m - i - 2;
        }
        for (int t = 0; t < 2; t++) {
            const int step = t ? step2 : step1;
            for (int c = 0; c < 256; c++) table[256 * t + c] = step + 1;
            for (int i = 0; i < m; i++) table[256 * t + x[i]] = step - i;
        }
        qq = step1;
        sh1 = step2;
        ntable = 512;
        break;
    }
    case 638: {  // qlqs.c:45-48
        static thread_local unsigned char rev[SB_XSIZE];
        for (int i = 0; i < m; i++) rev[i] = x[m - 1 - i];
        sb_prep_qsbc(x, m, table);
        sb_prep_qsbc(rev, m, table + 256);
        ntable = 512;
        break;
    }
    case 639:
    case 640: {  // tsw.c:52-53, tvsbs-w2.c:44-47
        static thread_local unsigned char rev[SB_XSIZE];
        for (int i = 0; i < m; i++) rev[i] = x[m - 1 - i];
        sb_prep_brbc(x, m, table);
        sb_prep_brbc(rev, m, table + 65536);
        ntable = 131072;
        break;
    }
    case 634:
        ntable = 1;
        table[0] = 0;
        break;
    case 635:
        sb_prep_mp(x, m, table);
        sb_prep_kmp(x, m, table, table + m + 1);
        sb_prep_buckets(x, m, table + 2 * (m + 1), table + 2 * (m + 1) + 256);
        ntable = 2 * (m + 1) + 256 + m;
        break;
    case 631:
        qq = desc->min_m;  // the q-gram length is the family's smallest m (bndmq2.c:35)
        sb_prep_bndm(x, m, reinterpret_cast<unsigned int *>(table));
        break;
    case 632: {
        const int k = m < SB_WORD ? m : SB_WORD;
        sb_prep_so(x, m, reinterpret_cast<unsigned int *>(table));  // fndm.c:36-37 == the masks of so.c:27-38
        lim = ~0u << (k - 1);                                         // NEG0m1, fndm.c:39
        break;
    }
    case 645: {  // aoso2.c:50-62 over the first min(m, 32) characters (search_large :106-117)
        const int k = m < SB_WORD ? m : SB_WORD;
        qq = desc->min_m - 1;  // aosoQ: m <= Q -> -1 (aoso2.c:47)
        unsigned h = 0, mm = 0;
        for (int c = 0; c < 256; c++) table[c] = -1;
        for (int j = 0; j < qq; j++) {
            for (int i = 0; i < k / qq; i++) {
                table[x[i * qq + j]] = (int)((unsigned)table[x[i * qq + j]] & ~(1u << h));
                h++;
            }
            mm |= 1u << (h - 1);
        }
        lim = mm;
        break;
    }
    case 644: {
        int st = 0, len = 0;
        qq = desc->min_m;  // bsdm: 1, bsdmQ: Q
        sb_prep_bsdm(x, m, qq, table, &st, &len);
        sh1 = st;
        lim = (unsigned)len;
        ntable = 65536;
        break;
    }
    case 641:
        qq = desc->min_m;  // bxs: 1, bxsQ: Q
        sb_prep_bndm(x, m, reinterpret_cast<unsigned int *>(table));  // bxs.c:44-48 (the first min(m, 32) characters)
        break;
    case 643: {  // lbndm.c:53-61: group g (from the right) of k characters shares bit WORD-1-g... as the reference loops
        const int k = (m - 1) / SB_WORD + 1;
        for (int c = 0; c < 256; c++) table[c] = 0;
        for (int i = m / k, l = m; i > 0; i--, l -= k)
            for (int j = k; j > 0; j--) table[x[l - j]] = (int)((unsigned)table[x[l - j]] | (1u << (SB_WORD - i)));
        break;
    }
    case 642:
    case 633: {
        const int k = m < SB_WORD ? m : SB_WORD;
        sb_prep_bndm(x, m, reinterpret_cast<unsigned int *>(table));
        for (int c = 0; c < 256; c++) table[c] = (int)((unsigned)table[c] << (SB_WORD - k));  // sbndm.c:42
        sb_prep_mp(x, k, scratch);
        sh1 = k - scratch[k];  // restore[0] (sbndm.c:44-55): the smallest period of the (32-character) window
        if (kind == 642) {     // sbndm-bmh.c:41-42
            for (int c = 0; c < 256; c++) table[256 + c] = 2 * k;
            for (int i = 0; i < k; i++) table[256 + x[i]] = 2 * k - i - 1;
            ntable = 512;
        }
        break;
    }
    case 628:
        sb_prep_bmgs(x, m, scratch, table);
        sb_prep_ztbc(x, m, table + m);
        ntable = m + 65536;
        break;
    case 629:
        qq = desc->min_m;  // the q-gram length is the family's smallest m (skip3.c:31)
        sb_prep_skipq(x, m, qq, table + m, table);
        ntable = m + 65536;
        break;
    case 630: {
        static thread_local unsigned char F[65536];
        // characters per scan step: 1 for wfr, Q for wfrqQ and for wfrQ (whose registry Q is 0: wfr's table builder)
        qq = (strcmp(desc->name, "wfr") == 0 || strcmp(desc->name, "twfr") == 0) ? 1 : desc->min_m;
        sh1 = desc->name[0] == 't';  // twfr.c:57, twfr4.c:63-66: the tuned variants skip windows whose last gram is no factor
        sb_prep_wfr(x, qq > 1 ? m - m % qq : m, desc->Q, F);  // wfrq4.c:75, wfr4.c:70: built over the cut pattern
        for (int i = 0; i < 2048; i++) {
            unsigned w = 0;
            for (int b = 0; b < 32; b++) w |= (unsigned)(F[i * 32 + b] != 0) << b;
            table[i] = (int)w;
        }
        ntable = 2048;
        break;
    }
    case 627: {
        sb_prep_mp(x, m, scratch);
        sb_prep_kmp(x, m, scratch, table);
        ntable = m + 1;
        int ell = 1;  // ac.c:52-53: the leading run of x[0], 0 when the pattern is one run
        while (ell < m && x[ell] == x[0]) ell++;
        qq = ell == m ? 0 : ell;
        break;
    }
    case 618:
    case 619:
        sb_prep_brbc(x, m, table);
        ntable = 65536;
        break;
    case 620:
        sb_prep_hbc(x, m, table);
        sb_prep_qsbc(x, m, table + 256);
        ntable = 512;
        break;
    case 621:
        sb_prep_buckets(x, m, table, table + 256);
        ntable = 256 + m;
        break;
    case 622:
        sb_prep_qsbc(x, m, table);
        sb_prep_mp(x, m, scratch);
        sb_prep_kmp(x, m, scratch, table + 256);
        ntable = 256 + m + 1;
        break;
    default:  // qf<Q><S>
        sb_prep_qf(x, m, desc->Q, desc->S, table);
        ntable = 1 << (desc->Q * desc->S);
        qq = desc->Q;
        sh1 = desc->S;
        break;
    }
    const double t1 = now_ms();
    const long long total_starts = (long long)t->n - m + 1;
    unsigned long long total = 0;
    double kernel_ms = 0;
    int launched = 0;
    for (int g = 0; g < t->nshards && total_starts > 0; g++) {  // one shard after the other: a blocking, selectable mode
        const Shard &s = t->sh[g];
        DevCtx *c = get_ctx(s.dev);
        if (!c) return SB_ERR_CUDA;
        if (!c->d_table) CU_TRY(cudaMalloc(&c->d_table, kTableInts * sizeof(int32_t)), SB_ERR_NOMEM);
        long long hi = total_starts - (long long)s.base;
        if (hi > (long long)s.own) hi = (long long)s.own;
        if (hi <= 0) continue;
        const size_t pat_bytes = stage_pattern(*c, kSyncSlot, x, m);
        CU_TRY(cudaMemcpyAsync(c->pat_dev(kSyncSlot), c->pat_host(kSyncSlot), pat_bytes, cudaMemcpyHostToDevice, c->stream),
               SB_ERR_CUDA);
        CU_TRY(cudaMemcpyAsync(c->d_table, table, (size_t)ntable * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream),
               SB_ERR_CUDA);
        CU_TRY(cudaMemsetAsync(c->cnt_dev(kSyncSlot), 0, 4 * sizeof(unsigned long long), c->stream), SB_ERR_CUDA);
        sb::TableArgs a;
        memset(&a, 0, sizeof a);
        a.text = s.d_text;
        a.pat = reinterpret_cast<const uint8_t *>(c->pat_dev(kSyncSlot));
        a.table = c->d_table;
        a.counters = c->cnt_dev(kSyncSlot);
        a.hi = hi;
        a.m = m;
        // start positions per thread; each thread also reads the m-1 bytes behind them.  4 KiB / 16 KiB keep that
        // halo small on a text that fills the device; a short text (the corpus buffers are 1-6 MiB) is cut finer so
        // that there are threads for every SM and a thread's sequential walk is short: down to 64 positions
        a.seg = m <= 1024 ? 4096 : 16384;
        {
            const long long want_threads = 148ll * 4 * sb::kTableThreads;
            long long fine = (hi / want_threads + 63) / 64 * 64;
            if (fine < 64) fine = 64;
            if (fine < a.seg) a.seg = (int)fine;
            const char *e = getenv("SMART_B200_TABLE_SEG");  // tests / tuning: force the segment length (multiple of 64)
            if (e && atoi(e) >= 64 && atoi(e) % 64 == 0) a.seg = atoi(e);
        }
        a.q = qq;
        a.sh1 = sh1;
        a.lim = lim;
        const unsigned long long nthreads = ((unsigned long long)hi + a.seg - 1) / a.seg;
        const unsigned grid = (unsigned)((nthreads + sb::kTableThreads - 1) / sb::kTableThreads);
        CU_TRY(cudaEventRecord(c->e0, c->stream), SB_ERR_CUDA);
        switch (kind) {
        case 601:
        case 602: sb::table_kmp_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 603: sb::table_hor_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 604:
            if (m > SB_WORD) sb::table_so_kernel<false, true><<<grid, sb::kTableThreads, 0, c->stream>>>(a);
            else sb::table_so_kernel<false, false><<<grid, sb::kTableThreads, 0, c->stream>>>(a);
            break;
        case 605:
            if (m > SB_WORD) sb::table_so_kernel<true, true><<<grid, sb::kTableThreads, 0, c->stream>>>(a);
            else sb::table_so_kernel<true, false><<<grid, sb::kTableThreads, 0, c->stream>>>(a);
            break;
        case 606: sb::table_hashq_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 607: sb::table_bm_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 608: sb::table_bc_kernel<sb::kBcQs><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 609: sb::table_bc_kernel<sb::kBcRaita><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 610: sb::table_bc_kernel<sb::kBcTuned><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 611:
            if (m <= SB_WORD) sb::table_kr_kernel<true><<<grid, sb::kTableThreads, 0, c->stream>>>(a);
            else sb::table_kr_kernel<false><<<grid, sb::kTableThreads, 0, c->stream>>>(a);
            break;
        case 613: {
            const size_t rows = (size_t)(m + 1) * 512;
            const size_t dyn = rows <= (size_t)sb::kAutSharedBytes ? rows : 0;
            CU_TRY(cudaFuncSetAttribute(sb::table_aut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        sb::kAutSharedBytes), SB_ERR_CUDA);
            sb::table_aut_kernel<<<grid, sb::kTableThreads, dyn, c->stream>>>(a);
            break;
        }
        case 614: sb::table_bom_kernel<false><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 636: sb::table_bom_kernel<true><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 615: sb::table_ebom_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 616: sb::table_sbndm2_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 617: sb::table_bndm_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 618: sb::table_br_kernel<false><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 619: sb::table_br_kernel<true><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 620: sb::table_smith_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 621: sb::table_skip_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 622: sb::table_fjs_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 623: sb::table_ssabs_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 624: sb::table_fs_kernel<false><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 648: sb::table_fs_kernel<true><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 646: sb::table_svm_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 649: sb::table_fbom_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 650:
        case 647: sb::table_wc_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 651: sb::table_pbmh_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 625: sb::table_graspm_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 626: sb::table_tbm_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 627: sb::table_ac_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 628: sb::table_zt_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 631:
            if (qq == 2) sb::table_bndmq_kernel<2><<<grid, sb::kTableThreads, 0, c->stream>>>(a);
            else if (qq == 4) sb::table_bndmq_kernel<4><<<grid, sb::kTableThreads, 0, c->stream>>>(a);
            else sb::table_bndmq_kernel<6><<<grid, sb::kTableThreads, 0, c->stream>>>(a);
            break;
        case 632: sb::table_fndm_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 634: sb::table_ts_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 637: sb::table_iom_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 638: sb::table_qlqs_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 639: sb::table_br2_kernel<false><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 640: sb::table_br2_kernel<true><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 635: sb::table_kmpskip_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 633: sb::table_sbndm_kernel<false><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 642: sb::table_sbndm_kernel<true><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 643: sb::table_lbndm_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 644: sb::table_bsdm_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 645: sb::table_aoso_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 641:
            switch (qq) {
            case 1: sb::table_bxs_kernel<1><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
            case 2: sb::table_bxs_kernel<2><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
            case 4: sb::table_bxs_kernel<4><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
            case 6: sb::table_bxs_kernel<6><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
            default: sb::table_bxs_kernel<8><<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
            }
            break;
        case 629: sb::table_skipq_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        case 630: sb::table_wfr_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        default: sb::table_qf_kernel<<<grid, sb::kTableThreads, 0, c->stream>>>(a); break;
        }
        CU_TRY(cudaGetLastError(), SB_ERR_CUDA);
        CU_TRY(cudaEventRecord(c->e1, c->stream), SB_ERR_CUDA);
        unsigned long long *h = c->cnt_host(kSyncSlot);
        CU_TRY(cudaMemcpyAsync(h, c->cnt_dev(kSyncSlot), 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream),
               SB_ERR_CUDA);
        CU_TRY(cudaStreamSynchronize(c->stream), SB_ERR_CUDA);
        float ms = 0;
        CU_TRY(cudaEventElapsedTime(&ms, c->e0, c->e1), SB_ERR_CUDA);
        kernel_ms += ms;  // shards are scanned one after the other
        total += h[0];
        launched++;
    }
    const double t2 = now_ms();
    g_timing.prep_host_ms = t1 - t0;
    g_timing.kernel_ms = kernel_ms;
    g_timing.wall_ms = t2 - t1;
    g_timing.kernels_launched = launched;
    g_timing.filter_kind = kind;
    if (pre_ms) *pre_ms = t1 - t0;
    if (run_ms) *run_ms = (t2 - t1) > 0 ? (t2 - t1) : 1e-6;
    return (long long)total;
}

long long sb_search64(int algo, const sb_text *t, const unsigned char *x, int m, unsigned flags,
                      unsigned long long *positions, size_t cap, double *pre_ms, double *run_ms)
{
    const sb_algo_desc *desc = sb_algo_desc_get(algo);
    if (!desc || !t || !x || m < 1 || m > SB_XSIZE) {
        set_err("bad argument to sb_search64 (algo/text/pattern/m)%s%s");
        return SB_ERR_ARG;
    }
    const bool want_pos = (flags & SB_FLAG_POSITIONS) != 0;
    if (want_pos && !positions && cap) {
        set_err("SB_FLAG_POSITIONS needs a positions buffer%s%s");
        return SB_ERR_ARG;
    }
    memset(&g_timing, 0, sizeof g_timing);
    if (pre_ms) *pre_ms = 0;
    if (run_ms) *run_ms = 0;
    if (want_pos && positions && cap > 0 && (long long)t->n - m + 1 > 0 && !outside_domain(desc, m, t->n))
        return search_positions(desc, t, x, m, flags, positions, cap, pre_ms, run_ms);
    if ((flags & SB_FLAG_TABLE_SCAN) && !outside_domain(desc, m, t->n)) {
        const long long r = search_table(desc, t, x, m, pre_ms, run_ms);
        if (r != -100) return r;
    }
    /* count only: one submit, one collect */
    const double t0 = now_ms();
    const long long idx = submit_core(desc, t, x, m, flags & ~SB_FLAG_POSITIONS);
    if (idx < 0) return idx;
    const double t2 = now_ms();
    const long long r = collect_core((int)idx);
    const double t3 = now_ms();
    g_timing.wall_ms = t3 - t2;
    if (r >= 0) {
        if (pre_ms) *pre_ms = t2 - t0;
        if (run_ms) *run_ms = (t3 - t2) > 0 ? (t3 - t2) : 1e-6;
    }
    return r;
}

/* Blocking search with position output: the min(count, cap) smallest positions, ascending.
 * Every shard is scanned ONCE, in windows of kPosWindow start positions: the scan kernels mark each occurrence in
 * the window's bitmap (emit_match), three small kernels compact the bitmap in order behind what the device has
 * already written (bitmap_count / bitmap_scan / bitmap_write) - no sort, and no second scan when `cap` is short:
 * the ordered write simply stops at cap while the count stays complete. */
static long long search_positions(const sb_algo_desc *desc, const sb_text *t, const unsigned char *x, int m, unsigned flags,
                                  unsigned long long *positions, size_t cap, double *pre_ms, double *run_ms)
{
    const double t0 = now_ms();
    Plan pl;
    plan_get(pl, desc, x, m, flags, t->bits, &t->pack, t->have_freq ? t->freq : nullptr, t->have_freq ? t->sig : 0);
    const double t1 = now_ms();
    const long long total_starts = (long long)t->n - m + 1;
    const size_t want = std::min((size_t)cap, (size_t)total_starts);

    DevCtx *ctx[kMaxShards];
    AuxSlot *aux[kMaxShards] = {nullptr};
    bool dev_first[kMaxShards];
    for (int g = 0; g < t->nshards; g++) {
        ctx[g] = get_ctx(t->sh[g].dev);
        if (!ctx[g]) return SB_ERR_CUDA;
        dev_first[g] = true;
        for (int k = 0; k < g; k++)
            if (t->sh[k].dev == t->sh[g].dev) {
                dev_first[g] = false;
                aux[g] = aux[k];
            }
        if (!dev_first[g]) continue;
        DevCtx &c = *ctx[g];
        if (!ensure_pos_capacity(c, want)) return SB_ERR_NOMEM;
        const size_t pat_bytes = stage_pattern(c, kSyncSlot, x, m);
        CU_TRY(cudaMemcpyAsync(c.pat_dev(kSyncSlot), c.pat_host(kSyncSlot), pat_bytes, cudaMemcpyHostToDevice, c.stream),
               SB_ERR_CUDA);
        if (pl.skip_grams) {
            aux[g] = aux_acquire(c, pl, c.stream);
            if (!aux[g]) return SB_ERR_CUDA;
        }
        CU_TRY(cudaMemsetAsync(c.cnt_dev(kSyncSlot), 0, 4 * sizeof(unsigned long long), c.stream), SB_ERR_CUDA);
        CU_TRY(cudaMemsetAsync(c.d_written, 0, sizeof(unsigned long long), c.stream), SB_ERR_CUDA);
        CU_TRY(cudaEventRecord(c.e0, c.stream), SB_ERR_CUDA);
    }
    const double t2 = now_ms();

    int launched = 0;
    const bool nothing = pl.filter == kFilterPacked && !pl.packed_K;  // a pattern byte the text lacks: count 0
    const int pbits = t->pack.packed ? t->pack.bits : 0;
    for (int g = 0; g < t->nshards && !nothing; g++) {  // shards in text order; each device's stream keeps that order
        DevCtx &c = *ctx[g];
        const Shard &s = t->sh[g];
        CU_TRY(cudaSetDevice(s.dev), SB_ERR_CUDA);
        const sb::ScanArgs whole = shard_args(pl, t, g, m);
        for (long long lo = 0; lo < whole.hi; lo += (long long)kPosWindow) {
            sb::ScanArgs a = whole;
            a.text = s.d_text + lo;  // a window is a shard of its own: its starts, the rest of the buffer as halo
            a.packed = (s.d_packed && pbits) ? s.d_packed + (unsigned long long)lo * pbits / 32 : nullptr;
            a.base = s.base + (unsigned long long)lo;
            a.hi = std::min<long long>(whole.hi - lo, (long long)kPosWindow);
            a.pat = c.pat_dev(kSyncSlot);
            a.counters = c.cnt_dev(kSyncSlot);
            a.aux = aux[g] ? aux[g]->d : nullptr;
            a.pos = c.d_posbits;
            const unsigned long long nwords = round_up(((unsigned long long)a.hi + 31) / 32, sb::kCompactWords);
            const unsigned nblocks = (unsigned)(nwords / sb::kCompactWords);
            CU_TRY(cudaMemsetAsync(c.d_posbits, 0, nwords * sizeof(uint32_t), c.stream), SB_ERR_CUDA);
            CU_TRY(launch_scan(c, pl, a, c.stream), SB_ERR_CUDA);
            launched++;
            sb::bitmap_count_kernel<<<nblocks, sb::kCompactThreads, 0, c.stream>>>(c.d_posbits, nwords, c.d_blocksum);
            sb::bitmap_scan_kernel<<<1, 1024, 0, c.stream>>>(c.d_blocksum, nblocks, c.d_blockoff, c.d_written);
            sb::bitmap_write_kernel<<<nblocks, sb::kCompactThreads, 0, c.stream>>>(c.d_posbits, nwords, c.d_blockoff, a.base, c.d_pos,
                                                                                   (unsigned long long)c.pos_cap < (unsigned long long)want
                                                                                       ? (unsigned long long)c.pos_cap
                                                                                       : (unsigned long long)want);
            CU_TRY(cudaGetLastError(), SB_ERR_CUDA);
        }
    }
    unsigned long long total = 0, verified = 0;
    double kernel_ms = 0;
    std::vector<unsigned long long> merged;  // only used when several devices contribute
    int ndev_used = 0;
    for (int g = 0; g < t->nshards; g++) ndev_used += dev_first[g] ? 1 : 0;
    size_t written = 0;
    for (int g = 0; g < t->nshards; g++) {
        if (!dev_first[g]) continue;
        DevCtx &c = *ctx[g];
        CU_TRY(cudaSetDevice(t->sh[g].dev), SB_ERR_CUDA);
        CU_TRY(cudaEventRecord(c.e1, c.stream), SB_ERR_CUDA);
        unsigned long long *h = c.cnt_host(kSyncSlot);
        CU_TRY(cudaMemcpyAsync(h, c.cnt_dev(kSyncSlot), 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream),
               SB_ERR_CUDA);
        CU_TRY(cudaMemcpyAsync(h + 1, c.d_written, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream), SB_ERR_CUDA);
        CU_TRY(cudaStreamSynchronize(c.stream), SB_ERR_CUDA);
        float ms = 0;
        CU_TRY(cudaEventElapsedTime(&ms, c.e0, c.e1), SB_ERR_CUDA);
        kernel_ms = std::max(kernel_ms, (double)ms);
        total += h[0];
        verified += h[2];
        if (h[1] != h[0]) {  // every counted occurrence must have been marked exactly once
            set_err("internal error: position bitmap and occurrence count disagree%s%s");
            return SB_ERR_CUDA;
        }
        const size_t take = (size_t)std::min<unsigned long long>(h[0], want);
        if (!take) continue;
        if (ndev_used == 1) {
            CU_TRY(cudaMemcpy(positions, c.d_pos, take * sizeof(unsigned long long), cudaMemcpyDeviceToHost), SB_ERR_CUDA);
            written = take;
        } else {
            const size_t old = merged.size();
            merged.resize(old + take);
            CU_TRY(cudaMemcpy(merged.data() + old, c.d_pos, take * sizeof(unsigned long long), cudaMemcpyDeviceToHost),
                   SB_ERR_CUDA);
        }
    }
    if (ndev_used > 1) {
        std::sort(merged.begin(), merged.end());  // each device's list is sorted; devices may interleave shards
        written = std::min(merged.size(), (size_t)cap);
        memcpy(positions, merged.data(), written * sizeof(unsigned long long));
    }
    (void)written;
    const double t3 = now_ms();

    g_timing.prep_host_ms = t1 - t0;
    g_timing.upload_ms = t2 - t1;
    g_timing.kernel_ms = kernel_ms;
    g_timing.wall_ms = t3 - t2;
    g_timing.kernels_launched = launched;
    g_timing.filter_kind = filter_kind_of(pl);
    g_timing.verified = verified;
    if (pre_ms) *pre_ms = t2 - t0;
    if (run_ms) *run_ms = (t3 - t2) > 0 ? (t3 - t2) : 1e-6;
    return (long long)total;
}


long long sb_search_multi64(int algo, const sb_text *t, const unsigned char *patterns, int npat, int m, unsigned flags,
                            unsigned long long *counts, double *pre_ms, double *run_ms)
{
    const sb_algo_desc *desc = sb_algo_desc_get(algo);
    if (!desc || !t || !patterns || !counts || npat < 1 || m < 1 || m > SB_XSIZE) {
        set_err("bad argument to sb_search_multi64%s%s");
        return SB_ERR_ARG;
    }
    memset(&g_timing, 0, sizeof g_timing);
    if (pre_ms) *pre_ms = 0;
    if (run_ms) *run_ms = 0;
    for (int p = 0; p < npat; p++) counts[p] = 0;
    if (outside_domain(desc, m, t->n)) return SB_NOT_APPLICABLE;

    /* block width: P(an aligned block equals one of the npat*B pattern grams) ~ npat*B*2^-(B*bits) must be
     * small; if no width achieves that (short patterns, tiny alphabets) fall back to one scan per pattern */
    int B = 0;
    /* preference 8, 16, 4: wider blocks mean fewer Bloom look-ups per text byte (scan_multi<16> runs at
     * 6.7 TB/s, <4> at 2.9 TB/s) but fewer patterns per pass (the exact table holds 4096 grams) */
    static const int order[3] = {8, 16, 4};
    for (int i = 0; i < 3; i++) {
        const int b = order[i];
        if (2 * b - 1 > m) continue;
        if (b * t->bits >= 14.0 + log2((double)npat * b)) {
            B = b;
            break;
        }
    }
    if (!B) {
        long long total = 0;
        double pre = 0, run = 0;
        for (int p = 0; p < npat; p++) {
            double a = 0, b = 0;
            long long c = sb_search64(algo, t, patterns + (size_t)p * m, m, flags & ~SB_FLAG_POSITIONS, nullptr, 0, &a, &b);
            if (c < 0) return c;
            counts[p] = (unsigned long long)c;
            total += c;
            pre += a;
            run += b;
        }
        if (pre_ms) *pre_ms = pre;
        if (run_ms) *run_ms = run > 0 ? run : 1e-6;
        g_timing.kernels_launched = npat * t->nshards;
        g_timing.filter_kind = -1;  // per-pattern fallback
        return total;
    }

    const double t0 = now_ms();
    const int nwords = (m + 3) / 4, stride_words = nwords + 1;
    const int anchor = desc->anchor == SB_ANCHOR_SUFFIX ? m - (2 * B - 1) : 0;
    const int batch_max = (sb::kSkipTableSize / 2) / B;
    const long long total_starts = (long long)t->n - m + 1;
    double prep_ms = 0, scan_ms = 0, kernel_ms = 0;
    unsigned long long verified = 0;
    int launched = 0;
    std::vector<uint32_t> hpat;
    for (int p0 = 0; p0 < npat; p0 += batch_max) {
        const int np = std::min(batch_max, npat - p0);
        const double b0 = now_ms();
        hpat.assign((size_t)np * stride_words, 0u);
        for (int p = 0; p < np; p++) memcpy(hpat.data() + (size_t)p * stride_words, patterns + (size_t)(p0 + p) * m, (size_t)m);
        // one staging per DEVICE
        bool dev_first[kMaxShards];
        DevCtx *ctx[kMaxShards];
        AuxSlot *maux[kMaxShards] = {nullptr};
        for (int g = 0; g < t->nshards; g++) {
            dev_first[g] = true;
            for (int k = 0; k < g; k++) dev_first[g] = dev_first[g] && (t->sh[k].dev != t->sh[g].dev);
            ctx[g] = get_ctx(t->sh[g].dev);
            if (!ctx[g]) return SB_ERR_CUDA;
            if (!dev_first[g]) continue;
            DevCtx &c = *ctx[g];
            if (c.mpat_words < hpat.size()) {
                CU_TRY(cudaStreamSynchronize(c.stream), SB_ERR_CUDA);
                if (c.d_mpat) cudaFree(c.d_mpat);
                c.d_mpat = nullptr;
                c.mpat_words = 0;
                if (cudaMalloc(&c.d_mpat, hpat.size() * sizeof(uint32_t)) != cudaSuccess) return SB_ERR_NOMEM;
                c.mpat_words = hpat.size();
            }
            if (c.mcounts_cap < (size_t)np) {
                CU_TRY(cudaStreamSynchronize(c.stream), SB_ERR_CUDA);
                if (c.d_mcounts) cudaFree(c.d_mcounts);
                c.d_mcounts = nullptr;
                c.mcounts_cap = 0;
                if (cudaMalloc(&c.d_mcounts, (size_t)np * sizeof(unsigned long long)) != cudaSuccess) return SB_ERR_NOMEM;
                c.mcounts_cap = (size_t)np;
            }
            // Bloom bitmap (3 probes) + exact table {fingerprint, ((pattern << 5) | d) + 1}
            AuxSlot *ax = aux_evict(c, c.stream);
            if (!ax) return SB_ERR_CUDA;
            maux[g] = ax;
            uint32_t *h_aux = ax->h;
            memset(h_aux, 0, kAuxWords * sizeof(uint32_t));
            uint32_t *table = h_aux + sb::kBloomWords;
            for (int p = 0; p < np; p++)
                for (int d = 0; d < B; d++) {
                    const uint32_t h = fingerprint(patterns + (size_t)(p0 + p) * m + anchor + d, B);
                    const uint32_t s1 = sb::bloom_slot(h), s2 = sb::bloom_slot2(h), s3 = sb::bloom_slot3(h);
                    h_aux[s1 >> 5] |= 1u << (s1 & 31);
                    h_aux[s2 >> 5] |= 1u << (s2 & 31);
                    h_aux[s3 >> 5] |= 1u << (s3 & 31);
                    uint32_t slot = sb::table_slot(h);
                    while (table[2 * slot + 1]) slot = (slot + 1) & (sb::kSkipTableSize - 1);
                    table[2 * slot] = h;
                    table[2 * slot + 1] = (((uint32_t)p << 5) | (uint32_t)d) + 1;
                }
            CU_TRY(cudaMemcpyAsync(ax->d, h_aux, kAuxWords * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream), SB_ERR_CUDA);
            CU_TRY(cudaMemcpyAsync(c.d_mpat, hpat.data(), hpat.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream),
                   SB_ERR_CUDA);
            CU_TRY(cudaMemsetAsync(c.d_mcounts, 0, (size_t)np * sizeof(unsigned long long), c.stream), SB_ERR_CUDA);
            CU_TRY(cudaMemsetAsync(c.cnt_dev(kSyncSlot), 0, 4 * sizeof(unsigned long long), c.stream), SB_ERR_CUDA);
            CU_TRY(cudaStreamSynchronize(c.stream), SB_ERR_CUDA);  // hpat is reused by the next device
        }
        const double b1 = now_ms();
        for (int g = 0; g < t->nshards; g++) {
            const Shard &sh = t->sh[g];
            DevCtx &c = *ctx[g];
            CU_TRY(cudaSetDevice(sh.dev), SB_ERR_CUDA);
            sb::ScanArgs a;
            memset(&a, 0, sizeof a);
            a.text = sh.d_text;
            a.pat = c.d_mpat;
            a.pat_stride_words = stride_words;
            a.npat = np;
            a.multi_counts = c.d_mcounts;
            a.counters = c.cnt_dev(kSyncSlot);
            for (int k = 0; k <= g; k++)  // the block staged by the first shard of this device
                if (maux[k] && t->sh[k].dev == sh.dev) a.aux = maux[k]->d;
            a.base = sh.base;
            a.m = m;
            a.nwords = nwords;
            a.tailmask = (m & 3) ? ((1u << (8 * (m & 3))) - 1u) : 0xffffffffu;
            a.anchor = anchor;
            long long hi = total_starts - (long long)sh.base;
            if (hi > (long long)sh.own) hi = (long long)sh.own;
            a.hi = hi > 0 ? hi : 0;
            if (dev_first[g]) CU_TRY(cudaEventRecord(c.e0, c.stream), SB_ERR_CUDA);
            if (a.hi > 0) {
                const unsigned long long end = (unsigned long long)a.hi + anchor + 2 * B;
                a.ngroups = (end + kGroupLong - 1) / kGroupLong;
                const int occ = c.occ_multi[B == 4 ? 0 : B == 8 ? 1 : 2];
                const unsigned long long want = (a.ngroups + (sb::kThreads / 32) - 1) / (sb::kThreads / 32);
                const unsigned grid = (unsigned)std::max(1ull, std::min(want, (unsigned long long)c.sms * occ));
                if (B == 4) sb::scan_multi_kernel<4, kU><<<grid, sb::kThreads, 0, c.stream>>>(a);
                else if (B == 8) sb::scan_multi_kernel<8, kU><<<grid, sb::kThreads, 0, c.stream>>>(a);
                else sb::scan_multi_kernel<16, kU><<<grid, sb::kThreads, 0, c.stream>>>(a);
                CU_TRY(cudaGetLastError(), SB_ERR_CUDA);
                launched++;
            }
        }
        std::vector<unsigned long long> part((size_t)np);
        for (int g = 0; g < t->nshards; g++) {
            bool last = true;
            for (int k = g + 1; k < t->nshards; k++) last = last && (t->sh[k].dev != t->sh[g].dev);
            if (!last) continue;
            DevCtx &c = *ctx[g];
            CU_TRY(cudaSetDevice(t->sh[g].dev), SB_ERR_CUDA);
            CU_TRY(cudaEventRecord(c.e1, c.stream), SB_ERR_CUDA);
            CU_TRY(cudaMemcpyAsync(part.data(), c.d_mcounts, (size_t)np * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                   c.stream),
                   SB_ERR_CUDA);
            CU_TRY(cudaMemcpyAsync(c.cnt_host(kSyncSlot), c.cnt_dev(kSyncSlot), 4 * sizeof(unsigned long long),
                                   cudaMemcpyDeviceToHost, c.stream),
                   SB_ERR_CUDA);
            CU_TRY(cudaStreamSynchronize(c.stream), SB_ERR_CUDA);
            float ms = 0;
            CU_TRY(cudaEventElapsedTime(&ms, c.e0, c.e1), SB_ERR_CUDA);
            kernel_ms = std::max(kernel_ms, (double)ms);
            verified += c.cnt_host(kSyncSlot)[2];
            for (int p = 0; p < np; p++) counts[p0 + p] += part[(size_t)p];
        }
        const double b2 = now_ms();
        prep_ms += b1 - b0;
        scan_ms += b2 - b1;
    }
    long long total = 0;
    for (int p = 0; p < npat; p++) total += (long long)counts[p];
    g_timing.prep_host_ms = prep_ms;
    g_timing.kernel_ms = kernel_ms;
    g_timing.wall_ms = now_ms() - t0;
    g_timing.kernels_launched = launched;
    g_timing.filter_kind = 200 + B;
    g_timing.verified = verified;
    if (pre_ms) *pre_ms = prep_ms;
    if (run_ms) *run_ms = scan_ms > 0 ? scan_ms : 1e-6;
    return total;
}

/* one pass of the host text through the device for up to kBatchJobs searches: every 64 MiB chunk is copied ONCE
 * and scanned by every job before the buffer is reused; the copy of chunk k+1 overlaps the scans of chunk k */
static long long host_batch_pass(DevCtx *c, sb_host_job *jobs, int njobs, const unsigned char *y, size_t n, double bits,
                                 const float *freq, uint64_t sig, int *launched_out, double *prep_out)
{
    const double t0 = now_ms();
    std::vector<Plan> plans((size_t)njobs);
    std::vector<AuxSlot *> aux((size_t)njobs, nullptr);
    std::vector<char> live((size_t)njobs, 0);
    int mmax = 1;
    for (int i = 0; i < njobs; i++) {
        sb_host_job &jb = jobs[i];
        const sb_algo_desc *desc = sb_algo_desc_get(jb.algo);
        jb.pre_ms = jb.run_ms = 0;
        if (!desc || !jb.x || jb.m < 1 || jb.m > SB_XSIZE) {
            jb.count = SB_ERR_ARG;
            continue;
        }
        if (outside_domain(desc, jb.m, n)) {
            jb.count = SB_NOT_APPLICABLE;
            continue;
        }
        jb.count = 0;
        if ((long long)n - jb.m + 1 <= 0) continue;
        plan_get(plans[i], desc, jb.x, jb.m, jb.flags & ~SB_FLAG_POSITIONS, bits, nullptr, freq, sig);
        live[i] = 1;
        mmax = std::max(mmax, jb.m);
        const size_t pat_bytes = stage_pattern(*c, kBatchSlot0 + i, jb.x, jb.m);
        CU_TRY(cudaMemcpyAsync(c->pat_dev(kBatchSlot0 + i), c->pat_host(kBatchSlot0 + i), pat_bytes, cudaMemcpyHostToDevice,
                               c->stream),
               SB_ERR_CUDA);
        if (plans[i].skip_grams) {
            aux[i] = aux_acquire(*c, plans[i], c->stream);
            if (!aux[i]) return SB_ERR_CUDA;
        }
    }
    CU_TRY(cudaMemsetAsync(c->cnt_dev(kBatchSlot0), 0, (size_t)njobs * 4 * sizeof(unsigned long long), c->stream), SB_ERR_CUDA);
    CU_TRY(cudaStreamSynchronize(c->stream), SB_ERR_CUDA);
    *prep_out = now_ms() - t0;

    /* chunk buffers: 64 MiB of start positions + halo each, double buffered */
    const size_t kChunk = (size_t)64 << 20;
    const size_t chunk = std::min(kChunk, round_up(n ? n : 1, kShardAlign));
    const size_t need = round_up(chunk + kHalo, sb::kPadBytes) + sb::kPadBytes;
    if (c->chunk_cap < need) {
        for (int b = 0; b < 2; b++) {
            if (c->d_chunk[b]) cudaFree(c->d_chunk[b] - sb::kPadBytes);
            c->d_chunk[b] = nullptr;
        }
        c->chunk_cap = 0;
        for (int b = 0; b < 2; b++) {
            uint8_t *blk = nullptr;
            if (cudaMalloc(&blk, need + 2 * sb::kPadBytes) != cudaSuccess) {
                set_err("cudaMalloc of a streaming chunk failed%s%s");
                return SB_ERR_NOMEM;
            }
            CU_TRY(cudaMemset(blk, 0, need + 2 * sb::kPadBytes), SB_ERR_CUDA);
            c->d_chunk[b] = blk + sb::kPadBytes;
        }
        c->chunk_cap = need;
    }
    int launched = 0;
    size_t k = 0;
    const long long starts_any = (long long)n;  // chunks are cut by start position; each job clips its own range
    for (size_t base = 0; (long long)base < starts_any; base += chunk, k++) {
        const int b = (int)(k & 1);
        const size_t len = std::min(n - base, chunk + (size_t)mmax - 1);
        CU_TRY(cudaMemcpyAsync(c->d_chunk[b], y + base, len, cudaMemcpyHostToDevice, c->cstream[b]), SB_ERR_CUDA);
        // bytes past `len` may be stale from an earlier chunk: zero what the last group can touch
        const size_t pad_end = std::min(c->chunk_cap, round_up(len, sb::kPadBytes) + sb::kPadBytes);
        CU_TRY(cudaMemsetAsync(c->d_chunk[b] + len, 0, pad_end - len, c->cstream[b]), SB_ERR_CUDA);
        for (int i = 0; i < njobs; i++) {
            if (!live[i]) continue;
            long long total_starts = (long long)n - jobs[i].m + 1;
            if (jobs[i].starts && (long long)jobs[i].starts < total_starts) total_starts = (long long)jobs[i].starts;
            const long long hi = std::min<long long>(total_starts - (long long)base, (long long)chunk);
            if (hi <= 0) continue;
            sb::ScanArgs a = plans[i].args;
            a.text = c->d_chunk[b];
            a.pat = c->pat_dev(kBatchSlot0 + i);
            a.counters = c->cnt_dev(kBatchSlot0 + i);
            a.aux = aux[i] ? aux[i]->d : nullptr;
            a.base = base;
            a.hi = hi;
            CU_TRY(launch_scan(*c, plans[i], a, c->cstream[b]), SB_ERR_CUDA);
            launched++;
        }
    }
    CU_TRY(cudaStreamSynchronize(c->cstream[0]), SB_ERR_CUDA);
    CU_TRY(cudaStreamSynchronize(c->cstream[1]), SB_ERR_CUDA);
    unsigned long long *h = c->cnt_host(kBatchSlot0);
    CU_TRY(cudaMemcpy(h, c->cnt_dev(kBatchSlot0), (size_t)njobs * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost),
           SB_ERR_CUDA);
    unsigned long long verified = 0;
    for (int i = 0; i < njobs; i++)
        if (live[i]) {
            jobs[i].count = (long long)h[4 * i];
            verified += h[4 * i + 2];
        }
    g_timing.verified = verified;
    g_timing.filter_kind = njobs == 1 && live[0] ? filter_kind_of(plans[0]) : 0;
    *launched_out = launched;
    return 0;
}

long long sb_search_host_batch(sb_host_job *jobs, int njobs, const unsigned char *y, size_t n)
{
    if (!jobs || njobs < 1 || (!y && n)) {
        set_err("bad argument to sb_search_host_batch%s%s");
        return SB_ERR_ARG;
    }
    memset(&g_timing, 0, sizeof g_timing);
    DevCtx *c = get_ctx(host_device());
    if (!c) return SB_ERR_CUDA;
    /* selectivity estimate and byte frequencies from a host-side sample of the text (first 64 KiB) */
    unsigned long long hist[256] = {0};
    for (size_t i = 0, e = std::min<size_t>(n, 65536); i < e; i++) hist[y[i]]++;
    float freq[256];
    uint64_t sig = 0;
    {
        double hd[256];
        for (int b = 0; b < 256; b++) hd[b] = (double)hist[b];
        freq_from_hist(hd, freq, &sig);
    }
    const bool stats = n >= 4096;
    const double bits = stats ? collision_bits(hist) : 8.0;
    long long total = 0;
    int launched = 0;
    double prep = 0;
    const double t0 = now_ms();
    for (int first = 0; first < njobs; first += kBatchJobs) {  // more jobs than slots: one pass of the text per group
        const int cnt = std::min(kBatchJobs, njobs - first);
        int l = 0;
        double p = 0;
        const long long rc = host_batch_pass(c, jobs + first, cnt, y, n, bits, stats ? freq : nullptr, stats ? sig : 0, &l, &p);
        if (rc < 0) return rc;
        launched += l;
        prep += p;
    }
    const double wall = now_ms() - t0;
    for (int i = 0; i < njobs; i++) {
        jobs[i].pre_ms = prep / njobs;
        jobs[i].run_ms = std::max((wall - prep) / njobs, 1e-6);
        if (jobs[i].count > 0) total += jobs[i].count;
    }
    g_timing.prep_host_ms = prep;
    g_timing.wall_ms = wall - prep;
    g_timing.kernels_launched = launched;
    return total;
}

long long sb_search_host(int algo, const unsigned char *x, int m, const unsigned char *y, size_t n, unsigned flags,
                         double *pre_ms, double *run_ms)
{
    if (pre_ms) *pre_ms = 0;
    if (run_ms) *run_ms = 0;
    if (!sb_algo_desc_get(algo) || !x || (!y && n) || m < 1 || m > SB_XSIZE) {
        set_err("bad argument to sb_search_host%s%s");
        return SB_ERR_ARG;
    }
    sb_host_job job;
    memset(&job, 0, sizeof job);
    job.algo = algo;
    job.x = x;
    job.m = m;
    job.flags = flags;
    const long long rc = sb_search_host_batch(&job, 1, y, n);
    if (rc < 0) return rc;
    if (job.count >= 0) {
        if (pre_ms) *pre_ms = job.pre_ms;
        if (run_ms) *run_ms = job.run_ms;
    }
    return job.count;
}

int sb_plugin_search(int algo, unsigned char *x, int m, unsigned char *y, int n, double *pre_ms, double *run_ms)
{
    if (n < 0) return -1;
    long long r = sb_search_host(algo, x, m, y, (size_t)n, 0, pre_ms, run_ms);
    if (r == SB_NOT_APPLICABLE) return -1;
    if (r < 0) {
        fprintf(stderr, "smart_b200: %s\n", sb_last_error());
        return -1;
    }
    return r > INT_MAX ? INT_MAX : (int)r;
}

}  // extern "C"
