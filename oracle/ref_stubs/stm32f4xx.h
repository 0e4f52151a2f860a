/* Minimal stand-in for the ST vendor header: lets the reference's keys_controlling.h (pulled in by
 * gps_master.c) preprocess on a host.  Only the one peripheral type that header names is declared.
 * TEST INFRASTRUCTURE ONLY (see ../gps_oracle.h). */
#ifndef GPSB_REF_STUB_STM32F4XX_H
#define GPSB_REF_STUB_STM32F4XX_H
typedef struct { int unused; } GPIO_TypeDef;
#endif
