/* oracle/port/port.h -- TEST INFRASTRUCTURE (see oracle/README.md). */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stddef.h>
#include <stdint.h>
#include <string.h>

uint32_t port_nxm_sad(const uint8_t* src, uint32_t src_stride, const uint8_t* ref, uint32_t ref_stride, uint32_t h,
                      uint32_t w);
void port_sad_loop(const uint8_t* src, uint32_t src_stride, const uint8_t* ref, uint32_t ref_stride, uint32_t bh,
                   uint32_t bw, uint64_t* best_sad, int16_t* xc, int16_t* yc, uint32_t ref_step, uint8_t skip,
                   int16_t sa_w, int16_t sa_h);
#endif
