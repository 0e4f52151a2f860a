/*
 * nav.c - navigation-bit stream from the sign of the prompt in-phase sum: 20-ms bit-edge
 * synchronisation from sign flips inside 4-ms slots, majority-vote data bits, preamble / polarity
 * detection, IS-GPS-200 parity and 10-word subframe assembly with a sub-bit subframe time stamp.
 *
 * Behaviour follows Firmware/project_main/GPS/nav_data.c (cited per function).  The reference keeps the
 * per-slot sample buffers in function statics shared by every channel (nav_data.c:48-51); here they
 * live in the gpsb_aux the caller passes (shared for the reference-named API, per channel in the
 * batched receiver).  A completed subframe is decoded into the channel's ephemeris container right away
 * (nav_data_decode.c:33, lc_decode_subframe), by this library and by the device-resident loop alike.
 *
 * The implementation lives in core/gpsb_loop_core.h (lc_nav_*), one source for this library and for the
 * device-resident tracking loop; this file binds it to the millisecond clock and the shared scratch.
 */
#include <stdlib.h>

#include <string.h>

#include "../../include/gpsb_flat_state.h"
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

/* nav_data_decode.c:33 (nav_data_decode.h) */
uint8_t gps_nav_data_decode_subframe(gps_ch_t* channel) { return channel ? lc_decode_subframe(channel) : 0; }

/* the word assembler fed bit by bit, ms counter advancing 20 per bit from ms0 (test and replay helper) */
void gpsb_host_feed_nav_bits(gps_ch_t* ch, const uint8_t* bits, uint32_t n, uint32_t ms0)
{
    if (!ch || !bits) return;
    for (uint32_t i = 0; i < n; i++) {
        gpsb_host_set_packet_cnt(ms0 + 20u * i);
        hx_nav_word_bit(ch, bits[i]);
    }
}

static uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
void gpsb_host_channel_eph(const gps_ch_t* ch, gpsb_flat_eph* o)
{
    const sdreph_t* d = &ch->eph_data;
    const eph_t* e = &d->eph;
    memset(o, 0, sizeof *o);
    o->sat = e->sat; o->iode = e->iode; o->iodc = e->iodc; o->sva = e->sva; o->svh = e->svh; o->week = e->week;
    o->code = e->code; o->flag = e->flag;
    o->toe_time = (int64_t)e->toe.time; o->toc_time = (int64_t)e->toc.time; o->ttr_time = (int64_t)e->ttr.time;
    o->toe_sec_bits = d2u(e->toe.sec); o->toc_sec_bits = d2u(e->toc.sec); o->ttr_sec_bits = d2u(e->ttr.sec);
    o->A = d2u(e->A); o->e = d2u(e->e); o->i0 = d2u(e->i0); o->OMG0 = d2u(e->OMG0); o->omg = d2u(e->omg);
    o->M0 = d2u(e->M0); o->deln = d2u(e->deln); o->OMGd = d2u(e->OMGd); o->idot = d2u(e->idot);
    o->crc = d2u(e->crc); o->crs = d2u(e->crs); o->cuc = d2u(e->cuc); o->cus = d2u(e->cus); o->cic = d2u(e->cic);
    o->cis = d2u(e->cis); o->toes = d2u(e->toes); o->fit = d2u(e->fit); o->f0 = d2u(e->f0); o->f1 = d2u(e->f1);
    o->f2 = d2u(e->f2);
    for (int i = 0; i < 4; i++) o->tgd[i] = d2u(e->tgd[i]);
    o->ctype = d->ctype; o->week_gpst = d->week_gpst; o->cnt = d->cnt; o->cntth = d->cntth; o->update = d->update;
    o->prn = d->prn; o->week_gst = d->week_gst; o->sub_cnt = d->sub_cnt; o->received_mask = d->received_mask;
    o->received_mask_proc = d->received_mask_proc; o->tow_gpst = d2u(d->tow_gpst);
}

/* observation pair of a channel as bit patterns, and a setter for the subframe time of week (test helpers) */
void gpsb_host_channel_obs(const gps_ch_t* ch, uint64_t out2[2])
{
    out2[0] = d2u(ch->obs_data.pseudorange_m);
    out2[1] = d2u(ch->obs_data.tow_s);
}
void gpsb_host_channel_set_tow(gps_ch_t* ch, double tow_gpst) { ch->eph_data.tow_gpst = tow_gpst; }

