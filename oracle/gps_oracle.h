/*
 * gps_oracle.h - CPU restatement of the reference's acquisition / E-P-L correlator arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import, link or
 * execute anything under oracle/.  The product library (stm32f4_sdr_gps_b200/) has no CPU fallback
 * and never calls into this file.
 *
 * Parity pinning: every function here is checked (tests/test_oracle_vs_reference.py) against the
 * UNMODIFIED reference C compiled from /root/reference into oracle/_ref/libgpsref.so, against the
 * known-answer values of BASELINE.md section 4 (reference simulator buffer, best_phase == 100, ...),
 * and against the committed fixtures under tests/golden/ that were generated from that reference
 * build (tests/golden/make_golden.py).  Nothing here is "parity unpinned".
 *
 * The restatement is written in the BYTE / BIT domain of SURVEY.md section 8(a) rather than in the
 * reference's pointer-walking form, so that it is an independent statement of the same arithmetic.
 * Citations are to Firmware/project_main/GPS/ ("PM/GPS/") of the reference.
 */
#ifndef GPS_ORACLE_H
#define GPS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_CHIPS        1023    /* PRN_LENGTH, PM/config.h:28            */
#define ORC_MS_BYTES     2046    /* PRN_SPI_WORDS_CNT*2, PM/config.h:27   */
#define ORC_MS_SAMPLES   16368   /* BITS_IN_PRN, PM/config.h:26           */
#define ORC_HALF_SUM     8184    /* BITS_IN_PRN/2, gps_misc.c:108,140     */
#define ORC_IF_HZ        4092000 /* IF_FREQ_HZ, PM/config.h:23            */

/* gps_generate_prn (gps_misc.c:317-372): C/A Gold code, one byte (0/1) per chip. prn 1..210. */
int orc_ca_code(int prn, uint8_t chips[ORC_CHIPS]);

/* gps_generate_prn_data2 (gps_misc.c:282-300): 16 samples per chip, shifted up by bits&15 samples,
 * no wrap-around.  rep must hold 2048 bytes; bytes 0..2045 are what the correlator reads. */
void orc_replica(const uint8_t chips[ORC_CHIPS], unsigned bits, uint8_t rep[2048]);

/* NCO word of gps_shift_to_zero_freq (gps_misc.c:219): (uint32_t)(freq_hz / 0.003810972f), fp32. */
uint32_t orc_nco_step(float freq_hz);
/* per-32-sample phase advance (gps_misc.c:220-221) */
uint32_t orc_nco_step32(uint32_t acc_step);

/* Carrier NCO + 1-bit mixer (gps_misc.c:211-240 / 244-274) from an explicit start phase.
 * Writes bytes 0..2043 of I and Q; bytes 2044..2045 are NOT written (reference loop bound 511 words).
 * Returns the accumulator after the 511 words. */
uint32_t orc_mix(const uint8_t sig[ORC_MS_BYTES], uint32_t acc0, uint32_t step32,
                 uint8_t* data_i, uint8_t* data_q);

/* gps_mult_and_summ (gps_misc.c:48-93): raw mismatch counts for one byte offset. */
void orc_corr_sums(const uint8_t* rep, const uint8_t* data_i, const uint8_t* data_q,
                   unsigned offset, int* sum_i, int* sum_q);

/* gps_correlation_iq (gps_misc.c:128-145) */
void orc_correlation_iq(const uint8_t* rep, const uint8_t* data_i, const uint8_t* data_q,
                        unsigned offset, int16_t* res_i, int16_t* res_q);
/* gps_correlation8 (gps_misc.c:98-122) */
int16_t orc_correlation8(const uint8_t* rep, const uint8_t* data_i, const uint8_t* data_q,
                         unsigned offset);
/* correlation_search (gps_misc.c:155-191) */
uint16_t orc_correlation_search(const uint8_t* rep, const uint8_t* data_i, const uint8_t* data_q,
                                unsigned start_shift, unsigned stop_shift,
                                uint16_t* aver_val, uint16_t* phase);

/* gps_rewind_if_phase (gps_misc.c:196-204): returns the new accumulator. */
uint32_t orc_rewind_if_phase(uint32_t accum, float if_freq_offset_hz, unsigned steps);

/* Fused cells, as the callers use the primitives ------------------------------------------- */

/* One acquisition / pre-track cell (acquisition.c:282-294, :198-209, tracking.c:403-426):
 * replica(bits) + stateless mix at freq_hz + search over [start,stop). */
uint16_t orc_search_cell(const uint8_t chips[ORC_CHIPS], const uint8_t sig[ORC_MS_BYTES],
                         float freq_hz, unsigned bits, unsigned start, unsigned stop,
                         uint16_t* aver_val, uint16_t* phase);

/* One tracking integrate-and-dump (tracking.c:115-138): offsets derived from code_phase_fine,
 * persistent NCO.  out6 = IE,QE,IP,QP,IL,QL.  Returns the accumulator after the step. */
uint32_t orc_track_epl(const uint8_t chips[ORC_CHIPS], const uint8_t sig[ORC_MS_BYTES],
                       float if_freq_offset_hz, uint32_t accum_in, float code_phase_fine,
                       int16_t out6[6]);

/* Explicit-parameter form matching the device request record (acc0, step32, offsets, bits). */
void orc_epl_explicit(const uint8_t chips[ORC_CHIPS], const uint8_t sig[ORC_MS_BYTES],
                      uint32_t acc0, uint32_t step32, unsigned off_e, unsigned off_p,
                      unsigned off_l, unsigned bits, int16_t out6[6]);

/* E/P/L byte offsets + sub-byte shift from code_phase_fine (tracking.c:115-130). */
void orc_epl_offsets(float code_phase_fine, unsigned* off_e, unsigned* off_p, unsigned* off_l,
                     unsigned* bits);

/* Timing loops for bench.py's "port" baseline (single thread). ------------------------------ */
/* n_ms x n_sv open-loop E/P/L steps on given per-(sv,ms) parameters; returns seconds. */
double orc_time_epl(const uint8_t* chips_all, unsigned n_sv, const uint8_t* signal, unsigned n_ms,
                    const uint32_t* acc0, const uint32_t* step32, const uint16_t* off_p,
                    const uint8_t* bits, int16_t* out);
/* n_sv x n_bins x n_ms full-window search cells; out[3*c] = max, phase, avr; returns seconds. */
double orc_time_sweep(const uint8_t* chips_all, unsigned n_sv, const uint8_t* signal, unsigned n_ms,
                      int first_bin_hz, int bin_step_hz, unsigned n_bins, unsigned bits,
                      uint16_t* out);

#ifdef __cplusplus
}
#endif
#endif
