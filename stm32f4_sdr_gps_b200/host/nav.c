/*
 * nav.c - navigation-bit stream from the sign of the prompt in-phase sum: 20-ms bit-edge
 * synchronisation from sign flips inside 4-ms slots, majority-vote data bits, preamble / polarity
 * detection, IS-GPS-200 parity and 10-word subframe assembly with a sub-bit subframe time stamp.
 *
 * Behaviour follows Firmware/project_main/GPS/nav_data.c (cited per function).  The reference keeps the
 * per-slot sample buffers in function statics shared by every channel (nav_data.c:48-51); here they
 * live in the gpsb_aux the caller passes (shared for the reference-named API, per channel in the
 * batched receiver).  Ephemeris field decoding (nav_data_decode.c) is outside the hot path: a completed
 * subframe is left in nav_data.subframe_data for whoever wants to decode it.
 *
 * The implementation lives in core/gpsb_loop_core.h (lc_nav_*), one source for this library and for the
 * device-resident tracking loop; this file binds it to the millisecond clock and the shared scratch.
 */
#include <stdlib.h>

#include "host_internal.h"

/* nav_data.c:257-352 */
void hx_nav_word_bit(gps_ch_t* ch, uint8_t new_bit) { lc_nav_word_bit(ch, new_bit, hx_now_ms()); }

void gps_nav_data_words_detection(gps_ch_t* channel, uint8_t new_bit) { if (channel) hx_nav_word_bit(channel, new_bit); }

/* nav_data.c:46-138 */
void hx_nav_new_code(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, int16_t new_i)
{
    if (lc_nav_new_code(ch, aux, index, new_i, hx_now_ms())) lc_refine_edge(ch, aux);
}

void gps_nav_data_analyse_new_code(gps_ch_t* channel, uint8_t index, int16_t new_i)
{
    if (channel) hx_nav_new_code(channel, &g_shared_aux, index, new_i);
}
