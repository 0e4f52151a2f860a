// Serial-latency microbenchmark of the pieces the control threads of k_track_run execute per millisecond
// (core/gpsb_loop_core.h): ONE thread runs each piece in a dependent chain and reports clock64 ticks per call.
// This is what bounds the device-resident loop - a single thread's dependent-issue latency, not throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -I include -o tools/bin/ubench_loop tools/ubench_loop.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "../stm32f4_sdr_gps_b200/core/gpsb_loop_core.h"

constexpr int N = 2000;

enum Piece { P_EMPTY, P_FDIV, P_DDIV_PI, P_ATANF, P_FLL_ANGLE, P_COSTAS_POS, P_COSTAS_NEG, P_FOLD, P_NCO, P_DLL, P_PLAN_CODE,
             P_F2D, P_LDS_CHAIN, P_CARRIER_FLL, P_CARRIER_PLL, P_CODE_STEP, P_NAV_STEP, P_LOCK_CHECK, P_PLAN_CARRIER, P_COUNT };
static const char* kNames[] = {"empty loop", "float divide (IEEE)", "float -> double, / pi, -> float", "lc_atanf", "lc_fll_angle (divide + atanf)",
                               "lc_costas_err, ip > 0 (atan2f float path)", "lc_costas_err, ip <= 0 (double atan2)", "lc_fold_half_pi",
                               "lc_nco_step32 (fadd, fdiv, f2u)", "lc_dll_update (state in shared memory)", "lc_arm_offsets",
                               "float -> double -> float", "dependent LDS",
                               "carrier thread step, slot index 1..3 (FLL)", "carrier thread step, slot index 0 (PLL)",
                               "code thread step (sums, DLL, offsets)", "nav thread step (nav bits, SNR)", "lc_lock_check (avg over a slot)",
                               "lc_plan_carrier"};

__global__ void k(int piece, long long* out, float* sink, int seed)
{
    __shared__ gps_ch_t ch;
    __shared__ gpsb_aux aux;
    __shared__ gpsb_epl_req srq;
    __shared__ uint32_t sums[4];
    __shared__ uint32_t chain[64];
    if (threadIdx.x != 0) return;
    for (int i = 0; i < 64; i++) chain[i] = (uint32_t)((i * 7 + 1) & 63);
    memset(&ch, 0, sizeof ch);
    memset(&aux, 0, sizeof aux);
    ch.tracking_data.code_phase_fine = 5000.0f;
    ch.tracking_data.state = GPS_TRACKING_RUN;
    ch.prn = 5;
    sums[0] = (8184u + 900u) | ((8184u - 300u) << 16);
    sums[1] = (8184u + 1800u + seed) | ((8184u - 500u) << 16);
    sums[2] = (8184u + 850u) | ((8184u - 280u) << 16);
    lc_angle_cache cache;
    cache.valid = 0;
    float x = 0.37f + seed * 1e-3f;
    int16_t ip = (int16_t)(3000 + seed), qp = (int16_t)(-1200 + seed);
    uint32_t u = seed;
    gpsb_epl_req rq;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) {
        switch (piece) {
        case P_EMPTY: x += 1.0f; break;
        case P_FDIV: x = 1.0f + x / 3.1f; break;
        case P_DDIV_PI: x = (float)((double)x / LC_PI) + 1.0f; break;
        case P_ATANF: x = lc_atanf(x) + 0.3f; break;
        case P_FLL_ANGLE: { float a = lc_fll_angle(ip, qp); qp = (int16_t)(qp + (lc_float_bits(a) & 3)); x += a; } break;
        case P_COSTAS_POS: { float a = lc_costas_err(ip, qp); qp = (int16_t)(qp + (lc_float_bits(a) & 3)); x += a; } break;
        case P_COSTAS_NEG: { float a = lc_costas_err((int16_t)-ip, qp); qp = (int16_t)(qp + (lc_float_bits(a) & 3)); x += a; } break;
        case P_FOLD: x = lc_fold_half_pi(x) + 0.9f; break;
        case P_NCO: { uint32_t s = lc_nco_step32(4092000.0f + x); x += (float)(s & 1u) + 0.5f; } break;
        case P_DLL: lc_dll_update(&ch.tracking_data, ip, qp, (int16_t)(ip - 5), (int16_t)(qp + (int16_t)ch.tracking_data.code_phase_fine % 3)); break;
        case P_PLAN_CODE: lc_arm_offsets(x, &rq); x += (float)(rq.off_p & 1u) + 0.25f; break;
        case P_F2D: { double d = (double)x; d += 1e-9; x = (float)d; } break;
        case P_LDS_CHAIN: u = ((volatile uint32_t*)chain)[u & 63]; break;
        case P_CARRIER_FLL:
        case P_CARRIER_PLL: {
            const uint8_t index = piece == P_CARRIER_PLL ? 0 : (uint8_t)(1 + i % 3);
            int16_t iq[6];
            const uint32_t packed[3] = {((volatile uint32_t*)sums)[0], ((volatile uint32_t*)sums)[1] + (uint32_t)(i & 7), ((volatile uint32_t*)sums)[2]};
            for (int a = 0; a < 3; a++) { iq[2 * a] = (int16_t)((int)(packed[a] & 0xFFFFu) - 8184); iq[2 * a + 1] = (int16_t)((int)(packed[a] >> 16) - 8184); }
            if (!lc_dll_is_degenerate(iq)) {
                lc_pll_update(&ch.tracking_data, ch.nav_data.period_sync_ok_flag, index, iq[2], iq[3]);
                lc_fll_update(&ch.tracking_data, &aux, ch.acq_data.found_freq_offset_hz, index, iq[2], iq[3], &cache);
                lc_plan_carrier(&ch.tracking_data, ch.prn, (uint32_t)i + 1, (uint32_t)i + 1, &srq);
            }
        } break;
        case P_CODE_STEP: {
            int16_t iq[6];
            const uint32_t packed[3] = {((volatile uint32_t*)sums)[0] + (uint32_t)(i & 7), ((volatile uint32_t*)sums)[1], ((volatile uint32_t*)sums)[2]};
            for (int a = 0; a < 3; a++) { iq[2 * a] = (int16_t)((int)(packed[a] & 0xFFFFu) - 8184); iq[2 * a + 1] = (int16_t)((int)(packed[a] >> 16) - 8184); }
            if (!lc_dll_is_degenerate(iq)) {
                lc_dll_update(&ch.tracking_data, iq[0], iq[1], iq[4], iq[5]);
                lc_plan_code(&ch.tracking_data, &srq);
            }
        } break;
        case P_NAV_STEP: {
            const int16_t ipv = (int16_t)(((i / 20) & 1) ? 1800 : -1800);
            if (lc_nav_new_code(&ch, &aux, (uint8_t)(i & 3), ipv, (uint32_t)i)) lc_refine_edge(&ch, &aux);
            lc_snr_update(&ch, &aux, ipv, (int16_t)200);
        } break;
        case P_LOCK_CHECK: lc_lock_check(&ch.tracking_data, &aux, 1000, (uint8_t)(i & 3), (int16_t)(1800 + (i & 1))); break;
        case P_PLAN_CARRIER: ch.tracking_data.if_freq_offset_hz += 0.25f; lc_plan_carrier(&ch.tracking_data, ch.prn, (uint32_t)i + 1, (uint32_t)i + 1, &srq); break;
        }
    }
    long long t1 = clock64();
    out[piece] = t1 - t0;
    sink[piece] = x + ch.tracking_data.code_phase_fine + ch.tracking_data.if_freq_offset_hz + (float)u + (float)qp + (float)srq.step32 + (float)srq.off_p;
}

int main()
{
    long long* d_out; float* d_sink;
    cudaMalloc(&d_out, P_COUNT * sizeof(long long));
    cudaMalloc(&d_sink, P_COUNT * sizeof(float));
    long long h[P_COUNT];
    for (int rep = 0; rep < 2; rep++)
        for (int p = 0; p < P_COUNT; p++) k<<<1, 32>>>(p, d_out, d_sink, rep + 1);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    printf("serial latency, one thread, clock64 ticks per call (loop overhead %.1f subtracted):\n", (double)h[P_EMPTY] / N);
    for (int p = 1; p < P_COUNT; p++) printf("  %-48s %8.1f\n", kNames[p], (double)(h[p] - h[P_EMPTY]) / N);
    return cudaGetLastError() != cudaSuccess;
}
