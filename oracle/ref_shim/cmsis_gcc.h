/* TEST INFRASTRUCTURE (oracle/_ref build only) -- not product code.
 *
 * Portable stand-in for <cmsis_gcc.h>, which the reference's affine.c /
 * affine_acc32b.c include unconditionally (ns-nnsp/src/affine.c:9). Only the four
 * Armv7E-M DSP-extension operations the reference uses are provided, written from
 * their architectural definitions (Arm ARM: ROR, SXTB16, SMLALD, SMLAD).
 */
#ifndef NNSP_ORACLE_CMSIS_GCC_SHIM_H
#define NNSP_ORACLE_CMSIS_GCC_SHIM_H
#include <stdint.h>

/* ROR: rotate a 32-bit word right by n (n in 1..31 at every reference call site) */
static inline uint32_t __ROR(uint32_t x, uint32_t n)
{
    n &= 31u;
    return n ? ((x >> n) | (x << (32u - n))) : x;
}

/* SXTB16: byte0 -> sign-extended low halfword, byte2 -> sign-extended high halfword */
static inline uint32_t __SXTB16(uint32_t x)
{
    uint32_t lo = (uint32_t)(uint16_t)(int16_t)(int8_t)(x & 0xffu);
    uint32_t hi = (uint32_t)(uint16_t)(int16_t)(int8_t)((x >> 16) & 0xffu);
    return lo | (hi << 16);
}

/* SMLALD: 64-bit acc += lo16(x)*lo16(y) + hi16(x)*hi16(y), all signed */
static inline uint64_t __SMLALD(uint32_t x, uint32_t y, uint64_t acc)
{
    int64_t p0 = (int64_t)(int16_t)(x & 0xffffu) * (int64_t)(int16_t)(y & 0xffffu);
    int64_t p1 = (int64_t)(int16_t)(x >> 16) * (int64_t)(int16_t)(y >> 16);
    return acc + (uint64_t)p0 + (uint64_t)p1;
}

/* SMLAD: 32-bit acc += lo*lo + hi*hi, result wraps modulo 2^32 (Q flag ignored) */
static inline uint32_t __SMLAD(uint32_t x, uint32_t y, uint32_t acc)
{
    uint32_t p0 = (uint32_t)((int32_t)(int16_t)(x & 0xffffu) * (int32_t)(int16_t)(y & 0xffffu));
    uint32_t p1 = (uint32_t)((int32_t)(int16_t)(x >> 16) * (int32_t)(int16_t)(y >> 16));
    return acc + p0 + p1;
}
#endif
