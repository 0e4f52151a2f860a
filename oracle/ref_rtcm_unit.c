/* TEST INFRASTRUCTURE ONLY (see gps_oracle.h).
 *
 * The reference ships with its RTCM output compiled out (config.h:30, ENABLE_RTCM_SEND 0) and guards the whole of
 * GPS/obs_publish.c with that switch.  This unit turns the switch on for that one file and compiles it from where it
 * lies, unmodified; config.h's include guard keeps the file's own #include "config.h" from turning it off again. */
#include "config.h"
#undef ENABLE_RTCM_SEND
#define ENABLE_RTCM_SEND 1
#include "GPS/obs_publish.c"
