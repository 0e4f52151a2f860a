/*
 * gpsb_kernels.cuh - device code of the B200 GPS L1 C/A correlator engine (sm_100a).
 *
 * Everything here is integer XOR / funnel-shift / popcount work on bit-packed 1-bit samples; there
 * is no floating point apart from the three IEEE-exact operations of the non-coherent detector
 * (reference PM/GPS/gps_misc.c:116-118).  Reference citations are to
 * Firmware/project_main/GPS/ of iliasam/STM32F4_SDR_GPS.
 *
 * Data model of one "cell" (one satellite x one millisecond x one NCO setting), all in shared memory:
 *
 *   R[512]    replica words: chip k covers sample bits [16k+b, 16k+b+16)      (gps_misc.c:282-300)
 *   I[1024]   mixed in-phase samples, extended periodically with period 2046 BYTES so that the
 *   Q[1024]   32-bit window at any byte position 4W+off (off <= 2045) is one funnel shift of two
 *             neighbouring words; bytes 2044..2045 of the period are zero because the reference
 *             mixer stops after 511 words                                     (gps_misc.c:229)
 *
 * The correlator for byte offset `off` is  sum_W popc((win(4W+off) ^ R[W]) & mask(W, off))  where
 * the mask reproduces the reference's word exclusions for odd offsets          (gps_misc.c:59-89).
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gpsb.h"

namespace gpsb {

constexpr int kWords = 512;        // 32-bit words per ms frame (incl. the 2 pad bytes)
constexpr int kMixWords = 511;     // PRN_SPI_WORDS_CNT/2, gps_misc.c:229
constexpr int kExtWords = 1024;    // two periods of 2046 bytes, rounded up
constexpr int kHalfSum = 8184;     // BITS_IN_PRN/2, gps_misc.c:108
constexpr int kEplThreads = 128;
constexpr int kSearchThreads = 256;

struct CellSmem {
    uint32_t R[kWords];
    uint32_t I[kExtWords];
    uint32_t Q[kExtWords];
};

// Quadrant patterns of the fs/4 carrier, gps_misc.c:216-217.  The reference literal 0x9999999 has
// seven nibbles, i.e. the top nibble is zero; sin[ph] == cos[(ph+3)&3].
__device__ __forceinline__ uint32_t cos_pattern(uint32_t ph)
{
    return (ph & 2u) ? ((ph & 1u) ? 0x33333333u : 0x66666666u)
                     : ((ph & 1u) ? 0xCCCCCCCCu : 0x09999999u);
}
__device__ __forceinline__ uint32_t sin_pattern(uint32_t ph) { return cos_pattern((ph + 3u) & 3u); }

// Replica word W for sub-byte shift b from the chip-expanded table E (E[w] = chips 2w, 2w+1 as
// 0xFFFF halves; E[511] = chip 1022 in the low half).  No wrap: bits below b stay 0.
__device__ __forceinline__ uint32_t replica_word(const uint32_t* __restrict__ E, int W, uint32_t b)
{
    uint32_t hi = __ldg(E + W);
    uint32_t lo = (W > 0) ? __ldg(E + W - 1) : 0u;
    return __funnelshift_l(lo, hi, b);
}

// Stage R for (slot table E, shift b).
__device__ __forceinline__ void stage_replica(uint32_t* R, const uint32_t* __restrict__ E, uint32_t b,
                                              int tid, int nthreads)
{
    for (int W = tid; W < kWords; W += nthreads) R[W] = replica_word(E, W, b);
}

// Carrier NCO + 1-bit mixer, closed form of gps_misc.c:229-239: the accumulator before word w is
// acc0 + w*step32 (mod 2^32).  Word 511 (bytes 2044..2047) is never mixed and stays 0.
__device__ __forceinline__ void stage_mix(uint32_t* I, uint32_t* Q, const uint32_t* __restrict__ frame,
                                          uint32_t acc0, uint32_t step32, int tid, int nthreads)
{
    for (int w = tid; w < kWords; w += nthreads) {
        uint32_t vi = 0u, vq = 0u;
        if (w < kMixWords) {
            uint32_t s = __ldcg(frame + w);   // L2 only: ring frames are rewritten by DMA under resident kernels
            uint32_t ph = (acc0 + (uint32_t)w * step32) >> 30;
            vi = cos_pattern(ph) ^ s;
            vq = sin_pattern(ph) ^ s;
        }
        I[w] = vi;
        Q[w] = vq;
    }
}

// Extend A[0..511] (one period = 2046 bytes, the upper half of A[511] is ignored) to 1024 words
// with period 2046 bytes.  Must be called by all threads of the CTA; contains two barriers.
template <int NT>
__device__ __forceinline__ void extend_period(uint32_t* A, uint32_t* B, int tid)
{
    constexpr int kPer = (kExtWords - kMixWords + NT - 1) / NT;  // words 511..1023
    uint32_t va[kPer], vb[kPer];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPer; i++) {
        int x = kMixWords + tid + i * NT;
        va[i] = vb[i] = 0u;
        if (x == kMixWords) {
            va[i] = (A[511] & 0xFFFFu) | (A[0] << 16);
            vb[i] = (B[511] & 0xFFFFu) | (B[0] << 16);
        } else if (x < kExtWords - 1) {
            int y = x - kWords;
            va[i] = __funnelshift_r(A[y], A[y + 1], 16);
            vb[i] = __funnelshift_r(B[y], B[y + 1], 16);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPer; i++) {
        int x = kMixWords + tid + i * NT;
        if (x < kExtWords) {
            A[x] = va[i];
            B[x] = vb[i];
        }
    }
    __syncthreads();
}

// Mask of replica word W for byte offset off (gps_misc.c:59-89): the upper half of word 511 is
// replica word 1023, which does not exist; for odd offsets 2k+1 the 16-bit replica words 1022-k and
// 1022 are skipped.
__device__ __forceinline__ uint32_t word_mask(int W, uint32_t off)
{
    uint32_t odd = off & 1u;
    if (W == kWords - 1) return odd ? 0u : 0x0000FFFFu;
    if (odd) {
        int u = 1022 - (int)(off >> 1);
        if (W == (u >> 1)) return (u & 1) ? 0x0000FFFFu : 0xFFFF0000u;
    }
    return 0xFFFFFFFFu;
}

// gps_correlation8's detector (gps_misc.c:108-118) on raw mismatch counts.
__device__ __forceinline__ int detector(int sum_i, int sum_q)
{
    int a = sum_i - kHalfSum, b = sum_q - kHalfSum;
    a = a < 0 ? 0 : a;
    b = b < 0 ? 0 : b;
    float s = __fadd_rn(__int2float_rn(a * a), __int2float_rn(b * b));
    return (int)(short)__float2int_rz(__fsqrt_rn(s));
}

// Raw mismatch counts of one offset over all 512 words (thread-serial; used by the search kernels).
__device__ __forceinline__ void corr_offset(const CellSmem& s, uint32_t off, int& sum_i, int& sum_q)
{
    const int x0 = (int)(off >> 2);
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t* __restrict__ pi = s.I + x0;
    const uint32_t* __restrict__ pq = s.Q + x0;
    uint32_t lo_i = pi[0], lo_q = pq[0];
    int si = 0, sq = 0;
#pragma unroll 8
    for (int W = 0; W < kWords - 1; W++) {
        uint32_t hi_i = pi[W + 1], hi_q = pq[W + 1];
        uint32_t r = s.R[W];
        si += __popc(__funnelshift_r(lo_i, hi_i, sh) ^ r);
        sq += __popc(__funnelshift_r(lo_q, hi_q, sh) ^ r);
        lo_i = hi_i;
        lo_q = hi_q;
    }
    {   // word 511: only replica word 1022 (low half), and only for even offsets
        uint32_t m = (off & 1u) ? 0u : 0x0000FFFFu;
        uint32_t r = s.R[kWords - 1];
        si += __popc((__funnelshift_r(lo_i, pi[kWords], sh) ^ r) & m);
        sq += __popc((__funnelshift_r(lo_q, pq[kWords], sh) ^ r) & m);
    }
    if (off & 1u) {  // take back the skipped replica word 1022-k (k = off>>1) unless it is word 1022
        int u = 1022 - (int)(off >> 1);
        int Wx = u >> 1;
        if (Wx != kWords - 1) {
            uint32_t r = s.R[Wx];
            uint32_t vi = __funnelshift_r(pi[Wx], pi[Wx + 1], sh) ^ r;
            uint32_t vq = __funnelshift_r(pq[Wx], pq[Wx + 1], sh) ^ r;
            uint32_t m = (u & 1) ? 0xFFFF0000u : 0x0000FFFFu;
            si -= __popc(vi & m);
            sq -= __popc(vq & m);
        }
    }
    sum_i = si;
    sum_q = sq;
}

}  // namespace gpsb
