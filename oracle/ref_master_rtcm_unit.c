/* TEST INFRASTRUCTURE ONLY (see gps_oracle.h).
 *
 * The reference ships with its RTCM output compiled out (config.h:30, ENABLE_RTCM_SEND 0).  With the switch on,
 * gps_master_nav_handling() calls gps_master_transmit_obs() ahead of gps_master_calculate_pos() on every idle slot
 * (gps_master.c:279-285), and that call refreshes the ONE observation array the sliced position solver reads between its
 * slices (gps_master.c:41, :439).  This unit compiles GPS/gps_master.c from where it lies, unmodified, with the switch
 * on - for _ref/libgpsref_rtcm.so, the variant of the compiled reference that tests/test_fix.py uses to pin the host
 * library's behaviour with gpsb_host_enable_rtcm(1).  config.h's include guard keeps the file's own #include "config.h"
 * from turning the switch off again. */
#include "config.h"
#undef ENABLE_RTCM_SEND
#define ENABLE_RTCM_SEND 1
#include "GPS/gps_master.c"
