/*
 * ingest.c - row N1 of SURVEY.md section 8(f), the file side: a raw recording on disk straight into a tracking run.
 *
 * The reference has no file input - its samples arrive over SPI into a DMA double buffer
 * (Firmware/project_main/signal_capture.c:9-16: 16-bit words, LSB first, one word per chip) - but recordings of that
 * bit stream are what PC_SpiLight (PC_SpiLight/Readme.txt:3-4, "-i rec_file.bin") produces and what a replay feeds.
 * Three containers are understood:
 *   - the MCU's own memory image (default): 2046 bytes per millisecond, sample n of a millisecond in bit n % 8 of byte
 *     n / 8 - the ring format; the file is mapped and streamed as it lies;
 *   - GPSB_FILE_MSB_FIRST: the same stream with the first sample of every byte in bit 7 (an SPI master reading MSB
 *     first, the MPSSE default); bytes are bit-reversed on the way in;
 *   - GPSB_FILE_IQ2: the MAX2769-native 2-bit I / 2-bit Q samples, one byte each (bit 0 = I sign), packed on the device
 *     (k_pack_iq2) behind the running loop.
 * The run itself is gpsb_rx_track_stream / _iq2: the device-resident loop starts on the first chunk and the rest of the
 * file is copied into the HBM ring while it tracks.
 */
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "host_internal.h"

static uint64_t bytes_per_ms(uint32_t flags) { return (flags & GPSB_FILE_IQ2) ? 16368u : 2046u; }

/* Whole milliseconds of signal in the file from first_byte on; negative gpsb_status when it cannot be read. */
int64_t gpsb_file_ms(const char* path, uint64_t first_byte, uint32_t flags)
{
    struct stat st;
    if (!path || stat(path, &st) != 0 || !S_ISREG(st.st_mode)) return hx_note(GPSB_ERR_ARG);
    if ((uint64_t)st.st_size < first_byte) return 0;
    return (int64_t)(((uint64_t)st.st_size - first_byte) / bytes_per_ms(flags));
}

static uint8_t reversed(uint8_t b)
{
    b = (uint8_t)((b >> 4) | (b << 4));
    b = (uint8_t)(((b & 0xCC) >> 2) | ((b & 0x33) << 2));
    return (uint8_t)(((b & 0xAA) >> 1) | ((b & 0x55) << 1));
}

/* Track milliseconds ms0 .. ms0 + n_ms - 1 from the recording in `path`, whose sample for ms0 starts at first_byte.
 * Logs as for gpsb_rx_track_run (may be NULL).  GPSB_ERR_ARG: no such file, or shorter than the run. */
int gpsb_rx_track_file(gpsb_rx* rx, const char* path, uint64_t first_byte, uint32_t ms0, uint32_t n_ms, uint32_t flags,
                       int16_t* iq_log, int8_t* nav_log)
{
    if (!rx || !path || n_ms == 0 || ((flags & GPSB_FILE_IQ2) && (flags & GPSB_FILE_MSB_FIRST))) return hx_note(GPSB_ERR_ARG);
    const int64_t have = gpsb_file_ms(path, first_byte, flags);
    if (have < 0) return (int)have;
    if ((uint64_t)have < n_ms) return hx_note(GPSB_ERR_ARG);

    const int fd = open(path, O_RDONLY);
    if (fd < 0) return hx_note(GPSB_ERR_ARG);
    const uint64_t span = (uint64_t)n_ms * bytes_per_ms(flags);
    const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE);
    const uint64_t map_from = first_byte / page * page, lead = first_byte - map_from;
    uint8_t* map = (uint8_t*)mmap(NULL, (size_t)(span + lead), PROT_READ, MAP_PRIVATE, fd, (off_t)map_from);
    close(fd);
    if (map == MAP_FAILED) return hx_note(errno == ENOMEM ? GPSB_ERR_NOMEM : GPSB_ERR_ARG);
    madvise(map, (size_t)(span + lead), MADV_SEQUENTIAL);
    const uint8_t* samples = map + lead;

    int rc;
    if (flags & GPSB_FILE_IQ2) {
        rc = gpsb_rx_track_stream_iq2(rx, ms0, n_ms, samples, 0, iq_log, nav_log);
    } else if (flags & GPSB_FILE_MSB_FIRST) {
        uint8_t table[256];
        for (int b = 0; b < 256; b++) table[b] = reversed((uint8_t)b);
        uint8_t* lsb = (uint8_t*)malloc((size_t)span);
        if (!lsb) { munmap(map, (size_t)(span + lead)); return hx_note(GPSB_ERR_NOMEM); }
        for (uint64_t k = 0; k < span; k++) lsb[k] = table[samples[k]];
        rc = gpsb_rx_track_stream(rx, ms0, n_ms, lsb, 0, iq_log, nav_log);
        free(lsb);
    } else {
        rc = gpsb_rx_track_stream(rx, ms0, n_ms, samples, 0, iq_log, nav_log);
    }
    munmap(map, (size_t)(span + lead));
    return rc;
}
