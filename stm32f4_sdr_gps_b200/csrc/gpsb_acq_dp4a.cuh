/*
 * gpsb_acq_dp4a.cuh - full-window acquisition search as byte-popcount correlation on the integer
 * dot-product pipe (IDP.4A), sm_100a.
 *
 * What is computed is exactly the reference's correlation_search() over gps_correlation8()
 * (Firmware/project_main/GPS/gps_misc.c:98-191) for a chip-expanded replica
 * (gps_generate_prn_data2, :282-300) on a mixed millisecond (gps_shift_to_zero_freq, :211-240).
 * How: the replica is constant over each chip (16 samples), so for a byte (half-chip) offset `off`
 *
 *     mismatches(off) = sum over chips c of  [ chip_c ? 16 - h : h ],   h = popcount of the 16 data
 *                                                                       bits under chip c
 *
 * With H[p] = popcount(data bits [8p+b, 8p+b+16)) for every byte position p (b = sub-byte replica
 * shift), the chips of an even / odd offset read H at even / odd positions only, so each parity is a
 * length-1023 circular cross-correlation of the +-1 chip sequence with a small-integer sequence:
 * 4 multiply-adds per dp4a instead of one XOR+POPC per 32 samples.  The pieces of the reference sum
 * that are not whole chips are restored bit-exactly in the epilogue from the mixed bit stream itself:
 *   - chip 1022 is cut to 16-b samples and the first b replica samples are 0 (no wrap, gps_misc.c:290-299)
 *   - odd offsets 2k+1 skip replica words 1022-k and 1022 (gps_misc.c:59-89)
 *   - data bytes 2044..2045 are never mixed and stay 0 (gps_misc.c:229)
 *
 * Default form (kHalf): one CTA of 256 threads = one mixed millisecond x up to NSV satellites x the 1023 offsets of ONE
 * parity; two CTAs per SM, so that while one is in its epilogue (the IEEE-sqrt detector of 4 x NSV results per thread,
 * an eighth of a CTA's time, during which the dot-product pipe would idle) the other is in its main loop.  The two
 * parities of a cell meet in a small global scratch record (atomic max / add); the CTA that arrives second writes the
 * reference's triple.  The undivided form is kept below for comparison:
 * One CTA = one mixed millisecond (ms, NCO phase/step, b) x up to NSV satellites x all 2046 offsets.
 * 512 threads; thread (parity, q) owns offsets 2*(4q+r)+parity, r = 0..3, for every satellite of the
 * tile: NSV x 4 x {I,Q} accumulators fed by two 16-byte shared loads (I and Q, four byte-shifted
 * copies of the H sequence interleaved) and NSV broadcast words of +-1 chips per step of 4 chips.
 */
#pragma once

#include "gpsb_kernels.cuh"

namespace gpsb {

constexpr int kAcqThreads = 512;
constexpr int kChipSteps = 256;        // 1024 chip slots / 4 per dp4a (chips 1022, 1023 are zero)
constexpr int kBitWords = 520;         // mixed bit stream, circularly extended past word 511
constexpr int kExtBytes = 2064;        // H sequence of one parity, two periods of 1023 + pad
constexpr int kXvLen = 516;

struct AcqGroup {                      // one mixed millisecond and the satellites searched on it
    uint32_t ms_index, acc0, step32;
    uint16_t off_bits, n_sv;
    uint32_t sv_slot[8];
    uint16_t start[8], stop[8];
    uint32_t res_index[8];
};

struct AcqScratch {                    // where the two parity halves of a cell meet (zeroed before the launch)
    unsigned key;
    int total;
    unsigned arrived;
    unsigned pad;
};

template <int NSV, bool kHalf = false>
struct AcqSmem {
    static constexpr int kPar = kHalf ? 1 : 2;   // parities handled by one CTA
    uint4 xv[2 * kPar][kXvLen];        // [iq*kPar+parity][a] = H bytes 4a+r .. 4a+r+3 for r = 0..3
    uint32_t S[kChipSteps][NSV];       // int8x4: +1 for chip 0, -1 for chip 1, 0 beyond chip 1021
    uint32_t bits[2][kBitWords];       // mixed I / Q sample bits
    uint8_t ext[2 * kPar][kExtBytes];  // H values, parity-split, two periods
    uint32_t chipbits[NSV][32];        // chips as a bitmap (bit c of the stream = chip c)
    int ones[NSV];                     // number of 1-chips among chips 0..1021
    unsigned key[NSV];
    int total[NSV];
};

__device__ __forceinline__ uint32_t win16(const uint32_t* __restrict__ bits, uint32_t bitpos)
{
    const uint32_t w = bitpos >> 5;
    return __funnelshift_r(bits[w], bits[w + 1], bitpos & 31u) & 0xFFFFu;
}

// 16-bit replica word u for sub-byte shift b from the chip bitmap (gps_misc.c:290-299).
__device__ __forceinline__ uint32_t replica16(const uint32_t* __restrict__ cb, int u, uint32_t b)
{
    const uint32_t cur = (cb[u >> 5] >> (u & 31)) & 1u ? 0xFFFFu : 0u;
    const uint32_t prev = (u > 0 && ((cb[(u - 1) >> 5] >> ((u - 1) & 31)) & 1u)) ? 0xFFFFu : 0u;
    return ((cur << b) | (prev >> (16u - b))) & 0xFFFFu;
}

template <int NSV, bool kSweep, bool kHalf = false>
__global__ void __launch_bounds__(kHalf ? kAcqThreads / 2 : kAcqThreads, kHalf ? 2 : 1)
k_acq_dp4a(const AcqGroup* __restrict__ groups, SweepParams sp, uint32_t n_sv_total,
           gpsb_search_res* __restrict__ res, const uint32_t* __restrict__ codes,
           const uint32_t* __restrict__ schips, const uint32_t* __restrict__ signal, uint32_t ring_ms,
           AcqScratch* __restrict__ scratch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AcqSmem<NSV, kHalf>& s = *reinterpret_cast<AcqSmem<NSV, kHalf>*>(smem_raw);
    constexpr int kThreads = kHalf ? kAcqThreads / 2 : kAcqThreads;
    constexpr int kPar = kHalf ? 1 : 2;
    const int tid = threadIdx.x;
    const uint32_t group_id = kHalf ? blockIdx.x >> 1 : blockIdx.x;
    const int my_par = kHalf ? (int)(blockIdx.x & 1u) : 0;     // the parity this CTA owns (kHalf)

    // ---- which cells
    __shared__ AcqGroup g;
    if (tid == 0) {
        if (kSweep) {   // group_id = (bin, ms) cell group; blockIdx.y = satellite tile
            const uint32_t gg = sp.group0 + group_id * (sp.group_skip + 1u);
            const uint32_t m = gg % sp.n_ms, b = gg / sp.n_ms;
            g.ms_index = sp.ms0 + m;
            g.acc0 = 0;
            g.step32 = sp.step32[b];
            g.off_bits = (uint16_t)sp.off_bits;
            const uint32_t v0 = blockIdx.y * NSV;
            const uint32_t n = n_sv_total - v0 < (uint32_t)NSV ? n_sv_total - v0 : (uint32_t)NSV;
            g.n_sv = (uint16_t)n;
            for (uint32_t v = 0; v < n; v++) {
                g.sv_slot[v] = sp.sv_slots[v0 + v];
                g.start[v] = 0;
                g.stop[v] = GPSB_OFFSETS;
                g.res_index[v] = sp.dense ? group_id * n_sv_total + (v0 + v) : ((v0 + v) * sp.n_bins + b) * sp.n_ms + m;
            }
        } else {
            g = groups[group_id];
        }
    }
    __syncthreads();
    const uint32_t bsh = g.off_bits & 15u;
    const int n_sv = g.n_sv;

    // ---- stage: mixed bit streams (closed-form NCO, gps_misc.c:229-239), chips, +-1 chip bytes
    {
        const uint32_t* frame = signal + (size_t)(g.ms_index % ring_ms) * kWords;
        for (int w = tid; w < kMixWords; w += kThreads) {
            const uint32_t sg = __ldg(frame + w);
            const uint32_t ph = (g.acc0 + (uint32_t)w * g.step32) >> 30;
            s.bits[0][w] = cos_pattern(ph) ^ sg;
            s.bits[1][w] = sin_pattern(ph) ^ sg;
        }
        for (int i = tid; i < kChipSteps * NSV; i += kThreads) {
            const int k4 = i / NSV, v = i % NSV;
            s.S[k4][v] = v < n_sv ? __ldg(schips + (size_t)g.sv_slot[v] * kChipSteps + k4) : 0u;
        }
        for (int i = tid; i < NSV * 32; i += kThreads) {
            const int v = i >> 5, w = i & 31;
            uint32_t m = 0;
            if (v < n_sv) {   // E[x] = chip 2x in the low half, chip 2x+1 in the high half
                const uint32_t* E = codes + (size_t)g.sv_slot[v] * kWords + w * 16;
#pragma unroll
                for (int x = 0; x < 16; x++) {
                    const uint32_t e = __ldg(E + x);
                    m |= (e & 1u) << (2 * x);
                    m |= ((e >> 16) & 1u) << (2 * x + 1);
                }
            }
            s.chipbits[v][w] = m;
        }
        if (tid < NSV) {
            s.key[tid] = 0;
            s.total[tid] = 0;
        }
    }
    __syncthreads();
    if (tid < 2 * (kBitWords - kMixWords)) {   // bytes 2044..2045 = 0, then the stream repeats from byte 0
        const int iq = tid / (kBitWords - kMixWords), x = kMixWords + tid % (kBitWords - kMixWords);
        const uint32_t* B = s.bits[iq];
        uint32_t v;
        if (x == kMixWords) v = B[0] << 16;
        else v = __funnelshift_r(B[x - kWords], B[x - kWords + 1], 16);
        s.bits[iq][x] = v;   // reads touch words 0..8 only, writes words 511..519: no hazard
    }
    if (tid < NSV) {
        int c = 0;
        for (int w = 0; w < 32; w++) {
            uint32_t m = s.chipbits[tid][w];
            if (w == 31) m &= 0x3FFFFFFFu;   // chips 992..1021 only
            c += __popc(m);
        }
        s.ones[tid] = c;
    }
    __syncthreads();

    // ---- H[p] = popcount of the 16 data bits under a chip starting at byte p (+ b bits)
    for (int i = tid; i < kPar * (int)GPSB_OFFSETS; i += kThreads) {
        // full form: every byte position p of both streams; half form: only the positions of this CTA's parity
        const int iq = i >= kPar * 1023, k = i - iq * kPar * 1023;
        const int p = kHalf ? 2 * k + my_par : k;
        const uint8_t h = (uint8_t)__popc(win16(s.bits[iq], 8u * p + bsh));
        uint8_t* e = s.ext[kHalf ? iq : iq * 2 + (p & 1)];
        e[p >> 1] = h;
        e[(p >> 1) + 1023] = h;
    }
    if (tid < 2 * kPar * (kExtBytes - 2046)) s.ext[tid / (kExtBytes - 2046)][2046 + tid % (kExtBytes - 2046)] = 0;
    __syncthreads();
    for (int i = tid; i < 2 * kPar * (kXvLen - 1); i += kThreads) {
        const int arr = i / (kXvLen - 1), a = i % (kXvLen - 1);
        const uint32_t* e32 = reinterpret_cast<const uint32_t*>(s.ext[arr]);
        const uint32_t lo = e32[a], hi = e32[a + 1];
        s.xv[arr][a] = make_uint4(lo, __funnelshift_r(lo, hi, 8), __funnelshift_r(lo, hi, 16),
                                  __funnelshift_r(lo, hi, 24));
    }
    __syncthreads();

    // ---- main loop: 4 chips per step
    const int q = tid & 255, par = kHalf ? my_par : tid >> 8;
    int accI[NSV][4], accQ[NSV][4];
#pragma unroll
    for (int v = 0; v < NSV; v++)
#pragma unroll
        for (int r = 0; r < 4; r++) accI[v][r] = accQ[v][r] = 0;
    {
        const uint4* __restrict__ xi = s.xv[kHalf ? 0 : par] + q;
        const uint4* __restrict__ xq = s.xv[kHalf ? 1 : 2 + par] + q;
#pragma unroll 2
        for (int k4 = 0; k4 < kChipSteps; k4++) {
            const uint4 hi = xi[k4], hq = xq[k4];
            uint32_t sv[NSV];
            if (NSV >= 4) {
#pragma unroll
                for (int v4 = 0; v4 < NSV / 4; v4++) {
                    const uint4 t = reinterpret_cast<const uint4*>(s.S[k4])[v4];
                    sv[4 * v4] = t.x; sv[4 * v4 + 1] = t.y; sv[4 * v4 + 2] = t.z; sv[4 * v4 + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int v = 0; v < NSV; v++) sv[v] = s.S[k4][v];
            }
#pragma unroll
            for (int v = 0; v < NSV; v++) {
                const int c = (int)sv[v];
                accI[v][0] = __dp4a((int)hi.x, c, accI[v][0]);
                accI[v][1] = __dp4a((int)hi.y, c, accI[v][1]);
                accI[v][2] = __dp4a((int)hi.z, c, accI[v][2]);
                accI[v][3] = __dp4a((int)hi.w, c, accI[v][3]);
                accQ[v][0] = __dp4a((int)hq.x, c, accQ[v][0]);
                accQ[v][1] = __dp4a((int)hq.y, c, accQ[v][1]);
                accQ[v][2] = __dp4a((int)hq.z, c, accQ[v][2]);
                accQ[v][3] = __dp4a((int)hq.w, c, accQ[v][3]);
            }
        }
    }

    // ---- epilogue: partial chip / skipped words, detector, first-max argmax
    unsigned key[NSV];
    int total[NSV];
#pragma unroll
    for (int v = 0; v < NSV; v++) { key[v] = 0; total[v] = 0; }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int j = 4 * q + r;
        if (j > 1022) continue;
        const uint32_t off = 2u * j + par;
        // samples of chip 1022 (16-b of them) followed by the b zero samples at the head of the replica
        uint32_t epos = 8u * off + 16352u + bsh;
        if (epos >= 16368u) epos -= 16368u;
        const uint32_t eI = win16(s.bits[0], epos), eQ = win16(s.bits[1], epos);
        uint32_t x1I = 0, x1Q = 0, x2I = 0, x2Q = 0;
        if (par) {   // odd offset 2k+1, k == j: words 1022-k (data byte 2045) and 1022 (data byte 2k-1) are skipped
            x1I = win16(s.bits[0], 8u * 2045u);
            x1Q = win16(s.bits[1], 8u * 2045u);
            if (j > 0) {
                x2I = win16(s.bits[0], 8u * (2u * j - 1u));
                x2Q = win16(s.bits[1], 8u * (2u * j - 1u));
            }
        }
#pragma unroll
        for (int v = 0; v < NSV; v++) {
            if (v >= n_sv) continue;
            const uint32_t* cb = s.chipbits[v];
            const uint32_t pat = ((cb[31] >> 30) & 1u) ? (0xFFFFu >> bsh) : 0u;   // chip 1022
            int cI = accI[v][r] + 16 * s.ones[v] + __popc(eI ^ pat);
            int cQ = accQ[v][r] + 16 * s.ones[v] + __popc(eQ ^ pat);
            if (par) {
                const uint32_t r1 = replica16(cb, 1022 - j, bsh);
                cI -= __popc(x1I ^ r1);
                cQ -= __popc(x1Q ^ r1);
                if (j > 0) {
                    const uint32_t r2 = replica16(cb, 1022, bsh);
                    cI -= __popc(x2I ^ r2);
                    cQ -= __popc(x2Q ^ r2);
                }
            }
            if (off >= g.start[v] && off < g.stop[v]) {
                const int c = detector(cI, cQ);
                total[v] += c;
                if (c > 0) {
                    const unsigned k = ((unsigned)c << 16) | (0xFFFFu - off);
                    key[v] = k > key[v] ? k : key[v];
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < NSV; v++) {
        if (v >= n_sv) continue;
        const unsigned k = __reduce_max_sync(0xFFFFFFFFu, key[v]);
        const int t = __reduce_add_sync(0xFFFFFFFFu, total[v]);
        if ((tid & 31) == 0) {
            atomicMax(&s.key[v], k);
            atomicAdd(&s.total[v], t);
        }
    }
    __syncthreads();
    if (tid < n_sv) {
        unsigned k = s.key[tid];
        int total = s.total[tid];
        if (kHalf) {
            // the other parity of this cell is another CTA's: meet in the scratch record, the second to arrive finishes
            AcqScratch* sc = scratch + g.res_index[tid];
            atomicMax(&sc->key, k);
            atomicAdd(&sc->total, total);
            __threadfence();
            if (atomicAdd(&sc->arrived, 1u) == 0u) return;
            __threadfence();
            k = atomicMax(&sc->key, 0u);
            total = atomicAdd(&sc->total, 0);
        }
        gpsb_search_res o;
        o.max = (uint16_t)(k >> 16);
        o.phase = k ? (uint16_t)(0xFFFFu - (k & 0xFFFFu)) : 0;
        o.avg = (uint16_t)(total / (2 * (int)GPSB_CHIPS));
        o.reserved = 0;
        res[g.res_index[tid]] = o;
    }
}

}  // namespace gpsb
